# single-GPU validation of the committed state: smoke(), GPU test suite, bench lines of the default workload and of cfg-1 / cfg-4
mkdir -p gpurun_out
TAG=${TAG:-r02H}
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/${TAG}_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/${TAG}_smoke.log
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/${TAG}_pytest_gpu.log
for wl in bbels_cifar10_k17 ls_mnist; do
  timeout 400 python bench.py --workload $wl --steps 10 --warmup 3 > gpurun_out/${TAG}_bench_$wl.json 2> gpurun_out/${TAG}_bench_$wl.err; echo "$wl bench rc=$?"; cut -c1-330 gpurun_out/${TAG}_bench_$wl.json
done
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/${TAG}_bench_1gpu.json 2> gpurun_out/${TAG}_bench_1gpu.err; echo "bench rc=$?"; cut -c1-330 gpurun_out/${TAG}_bench_1gpu.json
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${TAG}_bench_reference.json 2> gpurun_out/${TAG}_bench_reference.err; echo "reference arm rc=$?"; cut -c1-400 gpurun_out/${TAG}_bench_reference.json
