#!/bin/bash
# Scaling runs on one N-GPU box (from the repo root: gpurun --gpus N -- bash tools/scale_run.sh N [quick]).
# For the box's N: the driver's own flags (--steps 20 --warmup 5) and the long run (1000 steps) of the default workload
# -- they must agree -- at 1 and N ranks, the 256-stream strong-scaling config at 1 and N ranks, and the bare
# concurrent copies behind the e2e number.  JSON lines go to gpurun_out/scale_*.json.
set -u
N=${1:-8}
OUT=gpurun_out
mkdir -p $OUT
run() {  # n tag bench-args...
  local n=$1; shift; local tag=$1; shift
  if [ "$n" = 1 ]; then
    timeout 600 python bench.py --gpus 1 "$@" > $OUT/scale_${tag}.json 2> $OUT/scale_${tag}.err
  else
    timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 400)) \
      bench.py --gpus $n "$@" > $OUT/scale_${tag}.json 2> $OUT/scale_${tag}.err
  fi
  echo "== $tag rc=$?"; tail -c 300 $OUT/scale_${tag}.err | grep -v "^$" | tail -3
}
D="--steps 20 --warmup 5 --no-cpu-baseline"
L="--steps 1000 --warmup 20 --no-cpu-baseline --no-e2e --no-also"
S="--no-cpu-baseline --no-e2e --workload 720p_256streams_x16"
run 1 d1 $D
run $N d$N $D
run $N d${N}_again $D --no-e2e --no-also
run 1 l1 $L
run $N l$N $L
if [ "${2:-}" != quick ]; then
  run $N l${N}_nccl $L --exchange nccl
  run $N s$N --steps 100 --warmup 10 $S
  run 1 s1 --steps 50 --warmup 5 $S
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 400)) \
    tools/probe/link_concurrent.py > $OUT/link_concurrent_n$N.txt 2> $OUT/link_concurrent_n$N.err; echo "== link rc=$?"
  cat $OUT/link_concurrent_n$N.txt
fi
python - <<'PY'
import glob, json
for f in sorted(glob.glob("gpurun_out/scale_*.json")):
    for l in open(f).read().splitlines():
        if l.startswith("{"):
            d = json.loads(l)
            e = d.get("e2e", {})
            print(f.split("scale_")[1][:-5], "N", d["n_gpus"], "steps", d["steps"], round(d["value"]), "Mpx/s", round(d["ms_per_step"] * 1e3, 1), "us/step",
                  "kernel", round(d["roofline"]["kernel_us_single_launch_events"], 1), "per-rank ms", [round(x, 3) for x in d.get("per_rank_ms", {}).get("all", [])],
                  "xc", d.get("exchange_cost", {}), "e2e", round(e.get("value", 0)), "ceil", round(e.get("link_ceiling", {}).get("mpx_s_ceiling", 0)), d["clocks"]["reasons"],
                  {k: (round(v.get("value", 0)), round(v.get("efficiency_vs_single_gpu_same_run", 0), 3)) for k, v in (d.get("also") or {}).items()})
PY
