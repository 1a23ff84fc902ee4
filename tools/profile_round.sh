#!/bin/bash
# Round evidence on ONE GPU (from the repo root: gpurun -- bash tools/profile_round.sh r01):
#   1. bench.py (default flags) and bench.py --impl reference  -> gpurun_out/<tag>_bench.json, <tag>_bench_ref.json
#   2. every launch of a short bench run with its device time   -> gpurun_out/<tag>_launches.csv
#   3. ncu --set full of the fused kernel (2 launches)           -> gpurun_out/<tag>_tile_full.ncu-rep
# ncu runs only after the same command has exited 0 without it (B200_PROFILING.md).
TAG=${1:-r01}
OUT=gpurun_out
mkdir -p $OUT
python bench.py --impl reference --steps 50 --warmup 3 > $OUT/${TAG}_bench_ref.json 2> $OUT/${TAG}_bench_ref.err
python bench.py > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err; echo "bench rc=$?"
SHORT="python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu-baseline"
$SHORT > $OUT/plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file $OUT/${TAG}_launches.csv $SHORT > $OUT/ncu1.log 2>&1
echo "launch list rc=$?"
$SHORT > $OUT/plain2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:k_tile -s 4 -c 2 -f -o $OUT/${TAG}_tile_full $SHORT > $OUT/ncu2.log 2>&1
echo "full capture rc=$?"
tail -c 400 $OUT/${TAG}_bench.json
