#!/bin/bash
# round 2 diagnostics: where the mid-N sweep and the small-block step spend their time
OUT=gpurun_out; mkdir -p $OUT
export G6_B200_WAIT_SECONDS=120
echo "== launch list, N=131072 sweeps, K=32,8,0"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $OUT/launches_131k.csv python tools/block_stats.py --n 131072 --k 32,8,0 --abi-chunks 0 --sample 256 > $OUT/diag_131k.log 2>&1; tail -4 $OUT/diag_131k.log
echo "== launch list, N=16384 block steps"
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -s 400 -c 120 --csv --log-file $OUT/launches_lat16k.csv ./oracle/g6_latency amuse_b200/csrc/libsapporo.so 16384 60 > $OUT/diag_lat16k.log 2>&1; tail -9 $OUT/diag_lat16k.log
echo "== latency (no profiler)"; for n in 16384 131072; do timeout 120 ./oracle/g6_latency amuse_b200/csrc/libsapporo.so $n 2>&1 | tail -8; done | tee $OUT/latency_diag.log
echo "== K sweep normal lib"; timeout 300 python tools/block_stats.py --n 131072 --k 32,16,8,4,0 --abi-chunks 0 --sample 2048 2>&1 | tail -6 | tee $OUT/ksweep_131k.log
timeout 300 python tools/block_stats.py --n 16384 --k 32,16,8,4,0 --abi-chunks 1 --sample 2048 2>&1 | tail -11 | tee $OUT/ksweep_16k.log
