# single-GPU validation of the committed state: smoke(), GPU test suite, default bench line, cfg-4 bench line
mkdir -p gpurun_out
TAG=${TAG:-r02D}
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/${TAG}_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/${TAG}_smoke.log
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/${TAG}_pytest_gpu.log
timeout 300 python bench.py --workload bbels_cifar10_k17 --steps 10 --warmup 3 > gpurun_out/${TAG}_bench_bbels_cifar10_k17.json 2> gpurun_out/${TAG}_bench_bbels.err; echo "cfg4 bench rc=$?"; cut -c1-1200 gpurun_out/${TAG}_bench_bbels_cifar10_k17.json
if [ -z "$SKIP_HEADLINE" ]; then
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/${TAG}_bench_1gpu.json 2> gpurun_out/${TAG}_bench_1gpu.err; echo "bench rc=$?"; cut -c1-300 gpurun_out/${TAG}_bench_1gpu.json
fi
