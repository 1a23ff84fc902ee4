#!/bin/bash
# Scaling run on one 8-GPU box: weak scaling of the default workload at N = 1, 2, 4, 8 (peer-memory exchange), the same
# with the NCCL all-gather at N = 8 for comparison, and the strong-scaling 256-stream config at N = 8 and N = 1.
# Usage (from the repo root): gpurun --gpus 8 -- bash tools/scale_run.sh
set -u
OUT=gpurun_out
mkdir -p $OUT
run() {  # N extra-args tag
  local n=$1; shift; local tag=$1; shift
  if [ "$n" = 1 ]; then
    timeout 300 python bench.py --gpus 1 "$@" > $OUT/scale_${tag}.json 2> $OUT/scale_${tag}.err
  else
    timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 400)) \
      bench.py --gpus $n "$@" > $OUT/scale_${tag}.json 2> $OUT/scale_${tag}.err
  fi
  echo "== $tag rc=$?"; tail -c 600 $OUT/scale_${tag}.json | python -c "
import sys, json
for l in sys.stdin.read().splitlines():
    if l.startswith('{'):
        d = json.loads(l); print(d['n_gpus'], round(d['value']), 'Mpx/s', round(d['ms_per_step']*1e3, 1), 'us/step', 'e2e', round(d.get('e2e', {}).get('value', 0)), d['clocks'])
" 2>/dev/null || tail -3 $OUT/scale_${tag}.err
}
nvidia-smi topo -m > $OUT/topo_8gpu.txt 2>&1
python tools/probe/host_topo.py > $OUT/host_topo_8gpu.txt 2>&1
run 1 w1 --steps 1000 --warmup 20
run 2 w2 --steps 1000 --warmup 20 --no-cpu-baseline
run 4 w4 --steps 1000 --warmup 20 --no-cpu-baseline
run 8 w8 --steps 1000 --warmup 20 --no-cpu-baseline
run 8 w8_nccl --steps 1000 --warmup 20 --no-cpu-baseline --no-e2e --exchange nccl
run 8 w8_b --steps 1000 --warmup 20 --no-cpu-baseline --no-e2e
run 8 s8 --steps 200 --warmup 10 --no-cpu-baseline --no-e2e --workload 720p_256streams_x16
run 4 s4 --steps 100 --warmup 10 --no-cpu-baseline --no-e2e --workload 720p_256streams_x16
run 2 s2 --steps 100 --warmup 10 --no-cpu-baseline --no-e2e --workload 720p_256streams_x16
run 1 s1 --steps 50 --warmup 5 --no-cpu-baseline --no-e2e --workload 720p_256streams_x16
