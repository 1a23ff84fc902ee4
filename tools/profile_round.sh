#!/bin/bash
# Round evidence on ONE GPU (from the repo root: gpurun -- bash tools/profile_round.sh r02):
#   1. bench.py at the driver's flags, at the default flags, sustained (20 000 steps), and --impl reference
#   2. every launch of a short bench run with its device time                          -> <tag>_launches.csv
#   3. ncu --set full of the fused kernel on the three single-GPU BASELINE configs     -> <tag>_tile_full_<workload>.ncu-rep
#   4. single-pass counters of launches taken after ~2 s of continuous load            -> <tag>_sustained_launches.csv
#   5. the kernels beside the fused step                                               -> <tag>_aux_kernels.txt
# ncu runs only after the same command has exited 0 without it (B200_PROFILING.md).
TAG=${1:-r02}
OUT=gpurun_out
mkdir -p $OUT
python bench.py --impl reference --steps 20 --warmup 5 > $OUT/${TAG}_bench_ref.json 2> $OUT/${TAG}_bench_ref.err
python bench.py --steps 20 --warmup 5 > $OUT/${TAG}_bench_driver_flags.json 2> $OUT/${TAG}_bench_driver_flags.err; echo "bench (driver flags) rc=$?"
python bench.py > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err; echo "bench (default) rc=$?"
python bench.py --steps 20000 --warmup 20 --no-e2e --no-also --no-cpu-baseline > $OUT/${TAG}_bench_sustained.json 2> $OUT/${TAG}_bench_sustained.err; echo "bench (sustained) rc=$?"
SHORT="python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu-baseline --no-also"
$SHORT > $OUT/plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file $OUT/${TAG}_launches.csv $SHORT > $OUT/ncu1.log 2>&1
echo "launch list rc=$?"
for WL in 1080p_x64 4k_x32 720p_32streams_x16; do
  $SHORT --workload $WL > $OUT/plain2.log 2>&1 && \
  ncu --set full --clock-control none --import-source on -k regex:k_tile -s 4 -c 2 -f -o $OUT/${TAG}_tile_full_$WL $SHORT --workload $WL > $OUT/ncu2_$WL.log 2>&1
  echo "full capture $WL rc=$?"
done
LONG="python bench.py --steps 20000 --warmup 20 --no-e2e --no-cpu-baseline --no-also"
timeout 600 ncu --metrics gpu__time_duration.sum,sm__cycles_elapsed.avg.per_second,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum \
  --clock-control none -k regex:k_tile --launch-skip 18000 --launch-count 4 --csv --log-file $OUT/${TAG}_sustained_launches.csv $LONG > $OUT/ncu3.log 2>&1
echo "sustained launches rc=$?"
python tools/bench_aux.py > $OUT/${TAG}_aux_kernels.txt 2>&1; echo "aux rc=$?"
python tools/bench_aux.py 3840 2160 32 >> $OUT/${TAG}_aux_kernels.txt 2>&1
tail -c 300 $OUT/${TAG}_bench.json; echo; cat $OUT/${TAG}_aux_kernels.txt | head -12
