#!/bin/bash
# round 2: evidence that has to come back small: bench lines, ncu launch list, ncu --set full summaries as text
TAG=${1:-r02}
OUT=gpurun_out
mkdir -p $OUT
export G6_B200_WAIT_SECONDS=120
echo "== bench"; timeout 900 python bench.py --steps 3 --warmup 3 > $OUT/bench_n1_$TAG.log 2>&1; grep '^{' $OUT/bench_n1_$TAG.log > $OUT/bench_n1_$TAG.json; python -c "
import json; d=json.load(open('$OUT/bench_n1_$TAG.json')); print('value %.4g frac %.4f e2e %.4g pred %.3f ph4 %s parity %s' % (d['value'], d['roofline']['frac'], d['e2e']['value'], d['predictor']['frac'], d.get('ph4_s_per_unit',{}).get('value'), d['parity']['ok']))"
echo "== bench shuffled ids"; timeout 600 python bench.py --steps 2 --warmup 3 --shuffle-ids --no-e2e --no-cpu-baseline --no-extras 2>/dev/null | grep '^{' > $OUT/bench_shuffle_$TAG.json
echo "== reference arm"; timeout 600 python bench.py --impl reference --steps 2 --warmup 1 2>/dev/null | tail -1 > $OUT/bench_ref_$TAG.json
echo "== ncu launch list"; timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches_$TAG.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --no-extras --parity-sample 0 > $OUT/ncu_list_$TAG.log 2>&1
echo "== ncu full force"; timeout 900 ncu --set full --clock-control none --import-source on -k regex:force_fast -s 3 -c 1 -f -o /tmp/force_$TAG python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-extras --parity-sample 0 > $OUT/ncu_full_$TAG.log 2>&1
python tools/ncu_summary.py /tmp/force_$TAG.ncu-rep > $OUT/force_kernel_ncu_$TAG.txt 2>&1; head -40 $OUT/force_kernel_ncu_$TAG.txt
echo "== ncu full predict"; timeout 600 ncu --set full --clock-control none --import-source on -k regex:predict_kernel -s 2 -c 1 -f -o /tmp/predict_$TAG python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-extras --parity-sample 0 > $OUT/ncu_pred_$TAG.log 2>&1
python tools/ncu_summary.py /tmp/predict_$TAG.ncu-rep > $OUT/predict_kernel_ncu_$TAG.txt 2>&1; head -30 $OUT/predict_kernel_ncu_$TAG.txt
echo "== block stats 131k"; G6_B200_LIB=$PWD/amuse_b200/csrc/libsapporo_stats.so timeout 300 python tools/block_stats.py --n 131072 --k 32,0 --abi-chunks 2 --sample 512 2>&1 | tail -5 | tee $OUT/stats131k_$TAG.log
ls -la $OUT | tail -12
