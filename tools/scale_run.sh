#!/bin/bash
# Scaling run on one 8-GPU box (from the repo root: gpurun --gpus 8 -- bash tools/scale_run.sh [full]).
# Default: N = 1 and N = 8 of the default workload (weak) and of the 256-stream config (strong), peer-memory exchange,
# plus N = 8 with the NCCL all-gather.  "full" adds N = 2 and 4.
set -u
OUT=gpurun_out
mkdir -p $OUT
run() {  # N tag bench-args...
  local n=$1; shift; local tag=$1; shift
  if [ "$n" = 1 ]; then
    timeout 300 python bench.py --gpus 1 "$@" > $OUT/scale_${tag}.json 2> $OUT/scale_${tag}.err
  else
    timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 400)) \
      bench.py --gpus $n "$@" > $OUT/scale_${tag}.json 2> $OUT/scale_${tag}.err
  fi
  echo "== $tag rc=$?"
}
W="--steps 1000 --warmup 20 --no-cpu-baseline"
S="--no-cpu-baseline --no-e2e --workload 720p_256streams_x16"
run 1 w1 $W
if [ "${1:-}" = full ]; then run 2 w2 $W; run 4 w4 $W; fi
run 8 w8 $W
run 8 w8_nccl $W --no-e2e --exchange nccl
run 8 s8 --steps 200 --warmup 10 $S
if [ "${1:-}" = full ]; then run 4 s4 --steps 100 --warmup 10 $S; run 2 s2 --steps 100 --warmup 10 $S; fi
run 1 s1 --steps 50 --warmup 5 $S
python - <<'PY'
import glob, json
for f in sorted(glob.glob("gpurun_out/scale_*.json")):
    for l in open(f).read().splitlines():
        if l.startswith("{"):
            d = json.loads(l)
            print(f.split("scale_")[1][:-5], d["n_gpus"], round(d["value"]), "Mpx/s", round(d["ms_per_step"] * 1e3, 1), "us/step",
                  "e2e", round(d.get("e2e", {}).get("value", 0)), d["clocks"]["reasons"], d["config"].get("summary_exchange", "")[:11])
PY
