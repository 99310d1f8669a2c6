#!/bin/bash
# round 2 diagnostics: the small-block (masked kernel) step -- cost of the FP64 pairs, per-line profile
OUT=gpurun_out; mkdir -p $OUT
export G6_B200_WAIT_SECONDS=60
for k in 32 0; do echo "== latency K=$k"; G6_B200_KCLOSE=$k timeout 100 ./oracle/g6_latency amuse_b200/csrc/libsapporo.so 16384 2>&1 | tail -8; done | tee $OUT/latency_k.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:force_kernel -s 700 -c 1 -f -o /tmp/fk ./oracle/g6_latency amuse_b200/csrc/libsapporo.so 16384 60 > $OUT/diag3.log 2>&1
python tools/ncu_summary.py /tmp/fk.ncu-rep > $OUT/force_lat16k_ncu.txt 2>&1
ncu -i /tmp/fk.ncu-rep --page source --csv --print-source sass 2>/dev/null | gzip > $OUT/force_lat16k_sass.csv.gz
head -8 $OUT/force_lat16k_ncu.txt
