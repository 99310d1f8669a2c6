#!/bin/bash
# round 2: quick regression check after a kernel change: parity tests, mid-N K sweep, latency, headline
OUT=gpurun_out; mkdir -p $OUT
export G6_B200_WAIT_SECONDS=120
echo "== tests"; timeout 900 python -m pytest tests -m gpu -x -q -p xdist -n 1 2>&1 | tail -5
echo "== K sweep"; timeout 300 python tools/block_stats.py --n 131072 --k 32,16,0 --abi-chunks 2 --sample 1024 2>&1 | tail -6
timeout 300 python tools/block_stats.py --n 16384 --k 32,0 --abi-chunks 1 --sample 1024 2>&1 | tail -4
echo "== latency"; for n in 16384 131072; do timeout 120 ./oracle/g6_latency amuse_b200/csrc/libsapporo.so $n 2>&1 | tail -8; done
echo "== bench"; timeout 600 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $OUT/bench_check.log 2>&1; grep '^{' $OUT/bench_check.log > $OUT/bench_check.json; python -c "
import json; d=json.load(open('$OUT/bench_check.json')); print('value %.4g frac %.4f e2e %.4g pred %.3f ph4 %s parity %s' % (d['value'], d['roofline']['frac'], d['e2e']['value'], d['predictor']['frac'], d.get('ph4_s_per_unit',{}).get('value'), d['parity']))"
