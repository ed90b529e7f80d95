mkdir -p gpurun_out
for f in 0 1 2 4 3 5 6 7; do echo "== CDS_LS_DEBUG=$f"; CDS_LS_DEBUG=$f timeout 120 python tests/gpu_ls_umma_check.py 2>&1 | grep tcgen05; done > gpurun_out/ls_flags.log 2>&1
cat gpurun_out/ls_flags.log
