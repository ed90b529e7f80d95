# multi-GPU bench lines for profiles/: N=$1 ranks; workloads from $2... (default: headline, cfg-4)
N=$1; shift
mkdir -p gpurun_out
TAG=${TAG:-r02}
PORT=29600
for spec in "$@"; do
  name=${spec%%:*}; extra=${spec#*:}; [ "$extra" = "$spec" ] && extra=""
  PORT=$((PORT+1))
  out=gpurun_out/${TAG}_${name}_${N}gpu.json
  timeout ${TMO:-240} python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $PORT bench.py --gpus $N $extra 2> gpurun_out/${TAG}_${name}_${N}gpu.err | grep '"metric"' > $out
  echo "$name N=$N rc=${PIPESTATUS[0]}"; python -c "
import json,sys
try:
    d=json.load(open('$out')); print('  ms/step %.2f  pairs/s %.4e  e2e %.4e  clocks %s' % (d['ms_per_step'], d['value'], d['e2e']['value'], d['clocks'].get('per_rank_sm_mhz')))
except Exception as e: print('  no JSON line:', e)"
  tail -2 gpurun_out/${TAG}_${name}_${N}gpu.err | cut -c1-160
done
