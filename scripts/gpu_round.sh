#!/bin/bash
# One GPU session: parity tests, smoke, bench, ncu launch list + full capture of the force kernel.
# Usage (on the GPU box, from the repo root): bash scripts/gpu_round.sh [tag] [quick]
TAG=${1:-r01}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,power.limit --format=csv > $OUT/gpu_$TAG.txt 2>&1
echo "== pytest gpu" ; timeout 1800 python -m pytest tests -q -m gpu --timeout 900 -s > $OUT/pytest_$TAG.log 2>&1; tail -30 $OUT/pytest_$TAG.log
echo "== smoke" ; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5 | tee $OUT/smoke_$TAG.log
echo "== bench" ; timeout 900 python bench.py --steps 3 --warmup 3 2>&1 | tail -3 | tee $OUT/bench_$TAG.json
echo "== variants" ; for r in 1 0; do for v in 6 7 9 10; do G6_B200_REFINE=$r timeout 300 python bench.py --n 262144 --steps 2 --warmup 3 --variant $v --no-cpu-baseline --no-e2e 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('refine $r variant $v', '%.4g'%d['value'], '%.4f'%d['roofline']['frac'], d['roofline']['ms_per_launch'], d['clocks']['sm_mhz'], d['clocks']['reasons'])" ; done; done 2>&1 | tee $OUT/variants_$TAG.log
if [ "$2" != "quick" ]; then
echo "== ncu launch list" ; timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file $OUT/launches_$TAG.csv python bench.py --n 262144 --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > $OUT/ncu_list_$TAG.log 2>&1 ; tail -2 $OUT/ncu_list_$TAG.log
echo "== ncu full" ; timeout 900 ncu --set full --clock-control none --import-source on -k regex:force_ -s 4 -c 2 -f -o $OUT/force_$TAG python bench.py --n 262144 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > $OUT/ncu_full_$TAG.log 2>&1 ; tail -2 $OUT/ncu_full_$TAG.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:predict_kernel -s 1 -c 1 -f -o $OUT/predict_$TAG python bench.py --n 1048576 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > $OUT/ncu_pred_$TAG.log 2>&1 ; tail -2 $OUT/ncu_pred_$TAG.log
fi
ls -la $OUT
