#!/bin/bash
# round 2, first GPU session: parity tests, smoke, short bench
TAG=${1:-r02a}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,power.limit --format=csv > $OUT/gpu_$TAG.txt 2>&1
echo "== smoke" ; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5 | tee $OUT/smoke_$TAG.log
echo "== pytest gpu" ; timeout 1500 python -m pytest tests -q -m gpu --timeout 600 -s > $OUT/pytest_$TAG.log 2>&1; tail -40 $OUT/pytest_$TAG.log
echo "== bench 256k" ; timeout 300 python bench.py --n 262144 --steps 2 --warmup 3 --no-cpu-baseline --no-e2e 2>&1 | tail -2 | tee $OUT/bench256k_$TAG.json
echo "== bench 1M" ; timeout 600 python bench.py --steps 2 --warmup 3 --no-cpu-baseline 2>&1 | tail -2 | tee $OUT/bench_$TAG.json
ls -la $OUT
