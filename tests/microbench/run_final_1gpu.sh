# end-of-round single-GPU evidence: GPU test suite, the default bench line, the launch list and three single-launch ncu
# captures (kept small: gpurun copies back at most 64 MiB)
mkdir -p gpurun_out
TAG=${TAG:-r02}
if [ -z "$SKIP_TESTS" ]; then
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/${TAG}_pytest_gpu.log
fi
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/${TAG}_bench_1gpu.json 2> gpurun_out/${TAG}_bench_1gpu.err; echo "bench rc=$?"
timeout 300 python bench.py --steps 2 --warmup 3 --no-graph --no-cpu-baseline > gpurun_out/${TAG}_plain_nograph.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/${TAG}_launch_list.csv python bench.py --steps 2 --warmup 3 --no-graph --no-cpu-baseline > gpurun_out/${TAG}_ncu_list.log 2>&1; echo "launch list rc=$?"
export CDS_TRAJ_CACHE=gpurun_out/traj_b4.pt
for st in 5 10 17; do
  CDS_STEPS=$st python tests/gpu_step_profile.py > gpurun_out/${TAG}_plain_step$st.log 2>&1 && \
  CDS_STEPS=$st ncu --set full --clock-control none --import-source on -k regex:els_umma -s 2 -c 1 -o gpurun_out/${TAG}_els_umma_step$st python tests/gpu_step_profile.py > gpurun_out/${TAG}_ncu_step$st.log 2>&1; echo "ncu step $st rc=$?"
done
rm -f gpurun_out/traj_b4.pt
du -sh gpurun_out
