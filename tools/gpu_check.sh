#!/bin/bash
# One-GPU check (gpurun -- bash tools/gpu_check.sh TAG): the GPU test suite, smoke(), and bench.py at the driver's flags.
TAG=${1:-r02}
OUT=gpurun_out
mkdir -p $OUT
timeout 1500 python -m pytest tests -m gpu -x -q > $OUT/${TAG}_pytest.txt 2>&1; echo "pytest rc=$?"; tail -5 $OUT/${TAG}_pytest.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/${TAG}_smoke.txt 2>&1; echo "smoke rc=$?"; tail -2 $OUT/${TAG}_smoke.txt
timeout 600 python bench.py --steps 20 --warmup 5 > $OUT/${TAG}_bench_d.json 2> $OUT/${TAG}_bench_d.err; echo "bench rc=$?"; tail -c 600 $OUT/${TAG}_bench_d.err
head -c 3000 $OUT/${TAG}_bench_d.json
