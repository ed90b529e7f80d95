set -x
mkdir -p gpurun_out
export CDS_TRAJ_CACHE=gpurun_out/traj_b4.pt
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/r02c_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r02c_pytest_gpu.log
CDS_ELS_VARIANT=pv timeout 300 python tests/gpu_step_profile.py > gpurun_out/r02c_pv_plain.log 2>&1
export CDS_LIB_PATH=$PWD/convolutional_diffusion_b200/libcdscore_prof.so
export CDS_STEPS=3,6,8,10,12,14
for f in 0 2 8 16 24 32 18; do
  CDS_ELS_VARIANT=pv CDS_DEBUG_FLAGS=$f timeout 300 python tests/gpu_step_profile.py > gpurun_out/r02c_pv_flags_$f.log 2>&1
  echo "== flags $f"; grep -v "^#" gpurun_out/r02c_pv_flags_$f.log | cut -c1-60
done
for f in 0 2 8; do
  CDS_ELS_VARIANT=fma CDS_DEBUG_FLAGS=$f timeout 300 python tests/gpu_step_profile.py > gpurun_out/r02c_fma_flags_$f.log 2>&1
  echo "== fma flags $f"; grep -v "^#" gpurun_out/r02c_fma_flags_$f.log | cut -c1-60
done
