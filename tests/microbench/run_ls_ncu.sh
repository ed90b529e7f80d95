# one full ncu capture of the tensor-core LS kernel at cfg-1 (only after the identical command exited 0 without ncu)
mkdir -p gpurun_out
CMD="python bench.py --workload ls_mnist --steps 2 --warmup 3 --no-graph --no-cpu-baseline"
timeout 300 $CMD > gpurun_out/r02_ls_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:ls_umma -s 3 -c 1 -o gpurun_out/r02_ls_umma $CMD > gpurun_out/r02_ls_ncu_full.log 2>&1; echo "ncu full rc=$?"
du -sh gpurun_out
