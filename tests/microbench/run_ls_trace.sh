mkdir -p gpurun_out
CDS_LIB_PATH=$PWD/convolutional_diffusion_b200/libcdscore_prof.so CDS_LS_DEBUG=8 timeout 200 python tests/gpu_ls_clock_trace.py > gpurun_out/ls_trace.log 2>&1; echo rc=$?
sed -n 1,2p gpurun_out/ls_trace.log; sed -n 22,32p gpurun_out/ls_trace.log
