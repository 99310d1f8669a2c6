#!/bin/bash
# round 2: check after the masked-kernel change: tests, smoke, latency tables, phiGRAPE replay, bench line
TAG=${1:-r02b}
OUT=gpurun_out; mkdir -p $OUT
export G6_B200_WAIT_SECONDS=120
echo "== pytest -m gpu"; timeout 900 python -m pytest tests -m gpu -q -p xdist -n 1 2>&1 | tail -4 | tee $OUT/pytest_gpu_$TAG.log
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | tee $OUT/smoke_$TAG.log
echo "== latency"; for n in 1024 16384 131072; do timeout 120 ./oracle/g6_latency amuse_b200/csrc/libsapporo.so $n 300 2>&1 | tail -10; done | tee $OUT/latency_$TAG.txt
echo "== bench"; timeout 900 python bench.py --steps 3 --warmup 3 > $OUT/bench_n1_$TAG.log 2>&1; grep '^{' $OUT/bench_n1_$TAG.log > $OUT/bench_n1_$TAG.json; python -c "
import json; d=json.load(open('$OUT/bench_n1_$TAG.json')); print('value %.4g frac %.4f e2e %.4g pred %.3f ph4 %s parity %s' % (d['value'], d['roofline']['frac'], d['e2e']['value'], d['predictor']['frac'], d.get('ph4_s_per_unit',{}).get('value'), d['parity']['ok']))"
echo "== phigrape replay"; timeout 600 python tools/phigrape_timing.py 131072 1.0 1e-4 3000 2>&1 | tail -6 | tee $OUT/phigrape_$TAG.txt
