# A/B driver for the P.V epilogue (GPU box): full GPU test suite, per-step profile of both epilogues on a real trajectory,
# then the profiling switches of the profile build (CDS_DEBUG_FLAGS: 2 = no epilogue work, 16 = no P.V UMMAs, 32 = no MUFU)
mkdir -p gpurun_out
TAG=${TAG:-r02d}
export CDS_TRAJ_CACHE=gpurun_out/traj_b4.pt
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/${TAG}_pytest_gpu.log
for v in pv fma; do
  CDS_ELS_VARIANT=$v timeout 300 python tests/gpu_step_profile.py > gpurun_out/${TAG}_step_profile_$v.log 2>&1
  echo "== $v"; grep -v "^# i" gpurun_out/${TAG}_step_profile_$v.log | cut -c1-60
done
export CDS_LIB_PATH=$PWD/convolutional_diffusion_b200/libcdscore_prof.so
export CDS_STEPS=${STEPS:-3,6,8,10,12,14,16,19}
for f in ${FLAGS:-0 2 18}; do
  CDS_ELS_VARIANT=pv CDS_DEBUG_FLAGS=$f timeout 300 python tests/gpu_step_profile.py > gpurun_out/${TAG}_pv_flags_$f.log 2>&1
  echo "== pv flags $f"; grep -v "^#" gpurun_out/${TAG}_pv_flags_$f.log | cut -c1-220
done
