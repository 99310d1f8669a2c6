#!/bin/bash
# round 2, final single-GPU session: tests, smoke, bench (+ncu evidence), BASELINE config drivers
TAG=${1:-r02}
OUT=gpurun_out
mkdir -p $OUT
export G6_B200_WAIT_SECONDS=120
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,power.limit --format=csv > $OUT/gpu_$TAG.txt 2>&1
echo "== pytest gpu"; timeout 1200 python -m pytest tests -q --tb=short -p xdist -n 1 -m gpu 2>&1 | tail -8 | tee $OUT/pytest_$TAG.log
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | tee $OUT/smoke_$TAG.log
echo "== bench"; timeout 900 python bench.py --steps 3 --warmup 3 > $OUT/bench_n1_$TAG.log 2>&1; grep '^{' $OUT/bench_n1_$TAG.log > $OUT/bench_n1_$TAG.json; python -c "
import json; d=json.load(open('$OUT/bench_n1_$TAG.json')); print('value %.4g frac %.4f e2e %.4g pred %.3f ph4 %s parity %s' % (d['value'], d['roofline']['frac'], d['e2e']['value'], d['predictor']['frac'], d.get('ph4_s_per_unit',{}).get('value'), d['parity']['ok']))"
echo "== bench shuffled ids"; timeout 600 python bench.py --steps 2 --warmup 3 --shuffle-ids --no-e2e --no-cpu-baseline --no-extras > $OUT/bench_shuffle_$TAG.log 2>&1; grep '^{' $OUT/bench_shuffle_$TAG.log > $OUT/bench_shuffle_$TAG.json; python -c "
import json; d=json.load(open('$OUT/bench_shuffle_$TAG.json')); print('shuffled ids: value %.4g frac %.4f parity %s' % (d['value'], d['roofline']['frac'], d['parity']['ok']))"
echo "== reference arm"; timeout 600 python bench.py --impl reference --steps 2 --warmup 1 2>&1 | tail -1 | tee $OUT/bench_ref_$TAG.json | cut -c1-400
echo "== ncu launch list"; timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches_$TAG.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --no-extras --parity-sample 0 > $OUT/ncu_list_$TAG.log 2>&1; tail -1 $OUT/ncu_list_$TAG.log | cut -c1-200
echo "== ncu full force"; timeout 900 ncu --set full --clock-control none --import-source on -k regex:force_fast -s 3 -c 1 -f -o $OUT/force_$TAG python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-extras --parity-sample 0 > $OUT/ncu_full_$TAG.log 2>&1; tail -1 $OUT/ncu_full_$TAG.log | cut -c1-200
echo "== ncu full predict"; timeout 600 ncu --set full --clock-control none --import-source on -k regex:predict_kernel -s 2 -c 1 -f -o $OUT/predict_$TAG python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-extras --parity-sample 0 > $OUT/ncu_pred_$TAG.log 2>&1; tail -1 $OUT/ncu_pred_$TAG.log | cut -c1-200
echo "== sweep scaling"; timeout 900 python tools/sweep_scaling.py --gpus 1 --sizes 65536,131072,262144,524288,2097152,4194304 2>&1 | tee $OUT/sweep_scaling_1gpu_$TAG.txt
echo "== ph4 N=16k"; timeout 600 python tools/ph4_timing.py 16384 0.125 gpu 2>&1 | tail -2 | tee $OUT/ph4_16k_$TAG.txt
echo "== phigrape replay"; timeout 600 python tools/phigrape_timing.py 131072 1.0 1e-4 3000 2>&1 | tail -6 | tee $OUT/phigrape_$TAG.txt
ls -la $OUT | tail -30
