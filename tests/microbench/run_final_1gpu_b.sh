# end-of-round single-GPU evidence, part 2: default bench line, launch list + one full ncu capture of the cfg-4 (bbELS k=17)
# bench command (the tensor-core edge-band kernel); each ncu command only after the identical command exited 0 without ncu
mkdir -p gpurun_out
TAG=${TAG:-r02E}
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/${TAG}_bench_1gpu.json 2> gpurun_out/${TAG}_bench_1gpu.err; echo "bench rc=$?"; cut -c1-260 gpurun_out/${TAG}_bench_1gpu.json
CMD="python bench.py --workload bbels_cifar10_k17 --steps 2 --warmup 3 --no-graph --no-cpu-baseline"
timeout 300 $CMD > gpurun_out/${TAG}_cfg4_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_cfg4_launch_list.csv $CMD > gpurun_out/${TAG}_cfg4_ncu_list.log 2>&1; echo "launch list rc=$?"
timeout 300 $CMD > gpurun_out/${TAG}_cfg4_plain2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:bbels_edge_umma -s 3 -c 1 -o gpurun_out/${TAG}_bbels_edge_umma $CMD > gpurun_out/${TAG}_cfg4_ncu_full.log 2>&1; echo "ncu full rc=$?"
du -sh gpurun_out
