mkdir -p gpurun_out
timeout 400 python tests/gpu_ls_umma_check.py > gpurun_out/ls_check.log 2>&1; echo rc=$?; grep -c "used=True" gpurun_out/ls_check.log; tail -10 gpurun_out/ls_check.log
CDS_LIB_PATH=$PWD/convolutional_diffusion_b200/libcdscore_prof.so CDS_LS_DEBUG=8 timeout 200 python tests/gpu_ls_clock_trace.py > gpurun_out/ls_trace.log 2>&1; echo rc=$?
sed -n 1,2p gpurun_out/ls_trace.log; sed -n 20,30p gpurun_out/ls_trace.log
