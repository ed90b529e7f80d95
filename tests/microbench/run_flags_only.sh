# profiling switches only (profile build), selected steps of a cached trajectory
mkdir -p gpurun_out
TAG=${TAG:-r02f}
export CDS_TRAJ_CACHE=gpurun_out/traj_b4.pt
[ -f $CDS_TRAJ_CACHE ] || CDS_STEPS=3 python tests/gpu_step_profile.py > /dev/null 2>&1
export CDS_LIB_PATH=$PWD/convolutional_diffusion_b200/libcdscore_prof.so
export CDS_STEPS=${STEPS:-3,10,12}
for v in ${VARIANTS:-pv}; do
for f in ${FLAGS:-18 146 274 530 914}; do
  CDS_ELS_VARIANT=$v CDS_DEBUG_FLAGS=$f timeout 300 python tests/gpu_step_profile.py > gpurun_out/${TAG}_${v}_flags_$f.log 2>&1
  echo "== $v flags $f"; grep -v "^#" gpurun_out/${TAG}_${v}_flags_$f.log | cut -c1-230
done
done
