mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/gpu_dist_check.py > gpurun_out/r02_dist_check_2gpu.log 2>&1; echo "dist check rc=$?"; tail -3 gpurun_out/r02_dist_check_2gpu.log
TAG=r02 bash tests/microbench/run_scale.sh 2 "headline:--steps 10 --warmup 3"
