# A/B of product-type builds on selected steps of a cached trajectory: LIBS="a.so b.so" VARIANT=pv STEPS=3,10,12
mkdir -p gpurun_out
TAG=${TAG:-ab}
export CDS_TRAJ_CACHE=gpurun_out/traj_b4.pt
[ -f $CDS_TRAJ_CACHE ] || CDS_STEPS=3 python tests/gpu_step_profile.py > /dev/null 2>&1
export CDS_STEPS=${STEPS:-3,6,8,10,12,14}
for lib in $LIBS; do
  CDS_LIB_PATH=$PWD/convolutional_diffusion_b200/$lib CDS_ELS_VARIANT=${VARIANT:-pv} timeout 300 python tests/gpu_step_profile.py > gpurun_out/${TAG}_$lib.log 2>&1
  echo "== $lib"; grep -v "^#" gpurun_out/${TAG}_$lib.log | cut -c1-120
done
