#!/bin/bash
# round 2, final single-GPU evidence (small outputs only): tests, smoke, bench lines, ncu, block classes, BASELINE config drivers
TAG=${1:-r02}
OUT=gpurun_out
mkdir -p $OUT
export G6_B200_WAIT_SECONDS=120
echo "== pytest -m gpu"; timeout 900 python -m pytest tests -m gpu -q -p xdist -n 1 2>&1 | tail -4 | tee $OUT/pytest_gpu_$TAG.log
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | tee $OUT/smoke_$TAG.log
echo "== bench"; timeout 900 python bench.py --steps 3 --warmup 3 > $OUT/bench_n1_$TAG.log 2>&1; grep '^{' $OUT/bench_n1_$TAG.log > $OUT/bench_n1_$TAG.json; python -c "
import json; d=json.load(open('$OUT/bench_n1_$TAG.json')); print('value %.4g frac %.4f e2e %.4g pred %.3f ph4 %s parity %s' % (d['value'], d['roofline']['frac'], d['e2e']['value'], d['predictor']['frac'], d.get('ph4_s_per_unit',{}).get('value'), d['parity']['ok']))"
echo "== bench shuffled ids"; timeout 600 python bench.py --steps 2 --warmup 3 --shuffle-ids --no-e2e --no-cpu-baseline --no-extras 2>/dev/null | grep '^{' > $OUT/bench_shuffle_$TAG.json
echo "== bench K=0"; G6_B200_KCLOSE=0 timeout 600 python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-extras 2>/dev/null | grep '^{' > $OUT/bench_k0_$TAG.json
echo "== bench predictor A/B (4 CTAs/SM build)"; G6_B200_LIB=$PWD/amuse_b200/csrc/libsapporo_pb4.so timeout 600 python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-extras 2>/dev/null | grep '^{' > $OUT/bench_pb4_$TAG.json
python - <<PY
import json
for t in ("shuffle", "k0", "pb4"):
    try:
        d = json.load(open("$OUT/bench_%s_$TAG.json" % t)); print(t, "value %.4g frac %.4f pred %.3f jerk %.2e" % (d["value"], d["roofline"]["frac"], d["predictor"]["frac"], d["parity"]["device_path"]["jerk"]))
    except Exception as e:
        print(t, "failed", e)
PY
echo "== reference arm"; timeout 600 python bench.py --impl reference --steps 2 --warmup 1 2>/dev/null | tail -1 > $OUT/bench_ref_$TAG.json
echo "== ncu launch list"; timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches_$TAG.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --no-extras --parity-sample 0 > $OUT/ncu_list_$TAG.log 2>&1
python tools/launch_summary.py $OUT/launches_$TAG.csv | head -24 | tee $OUT/launch_summary_$TAG.txt
echo "== ncu full force"; timeout 900 ncu --set full --clock-control none --import-source on -k regex:force_fast -s 3 -c 1 -f -o /tmp/force_$TAG python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-extras --parity-sample 0 > $OUT/ncu_full_$TAG.log 2>&1
python tools/ncu_summary.py /tmp/force_$TAG.ncu-rep > $OUT/force_kernel_ncu_$TAG.txt 2>&1; grep -E "duration|fma_cycles_active|issue_active|dram__bytes|hot loop" $OUT/force_kernel_ncu_$TAG.txt
echo "== ncu full predict"; timeout 600 ncu --set full --clock-control none --import-source on -k regex:predict_kernel -s 2 -c 1 -f -o /tmp/predict_$TAG python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-extras --parity-sample 0 > $OUT/ncu_pred_$TAG.log 2>&1
python tools/ncu_summary.py /tmp/predict_$TAG.ncu-rep > $OUT/predict_kernel_ncu_$TAG.txt 2>&1; grep -E "duration|dram__bytes|warps_active" $OUT/predict_kernel_ncu_$TAG.txt
echo "== block classes"; for n in 1048576 131072 16384; do G6_B200_LIB=$PWD/amuse_b200/csrc/libsapporo_stats.so timeout 300 python tools/block_stats.py --n $n --k 32 --abi-chunks 0 --sample 512 2>&1 | tail -1; done | tee $OUT/block_classes_$TAG.txt
echo "== sweep scaling"; timeout 900 python tools/sweep_scaling.py --gpus 1 --sizes 65536,131072,262144,524288,2097152,4194304 2>&1 | tail -9 | tee $OUT/sweep_scaling_1gpu_$TAG.txt
echo "== latency"; for n in 1024 16384 131072; do timeout 120 ./oracle/g6_latency amuse_b200/csrc/libsapporo.so $n 300 2>&1 | tail -10; done | tee $OUT/latency_$TAG.txt
echo "== phigrape replay"; timeout 600 python tools/phigrape_timing.py 131072 1.0 1e-4 3000 2>&1 | tail -6 | tee $OUT/phigrape_$TAG.txt
ls -la $OUT | wc -l; du -sh $OUT
