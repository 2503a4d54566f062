# usage: bash tools/gpu_inflight_sweep.sh "<lib or ->  <max_blocks> <inflight>" ...
for cfg in "$@"; do
  set -- $cfg
  lib=$1; [ "$lib" = "-" ] && lib=""
  FMB_LIB=$lib FMB_MAX_BLOCKS=$2 python bench.py --inflight $3 --no-batch --no-3d --no-costmap --no-cpu-baseline --steps 12 --warmup 3 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('lib=$1 max_blocks=$2 inflight=$3', 'ms_per_step', round(d['ms_per_step'],3), 'kernel_ms', round(d['roofline']['kernel_ms'],3), 'lat', round(d.get('latency_ms_one_query',0),2), 'e2e_pipe', round(d['e2e_pipelined']['ms_per_step'],2), 'e2e', round(d['e2e']['ms_per_step'],2))
"
done
