#!/bin/bash
# round 2 diagnostics: per-line profile of the force kernel on a mid-N sweep (where the FP64-radius blocks cost most)
OUT=gpurun_out; mkdir -p $OUT
export G6_B200_WAIT_SECONDS=120
timeout 600 ncu --set full --clock-control none --import-source on -k regex:force_fast -s 1 -c 1 -f -o /tmp/f131 python tools/block_stats.py --n 131072 --k 32 --abi-chunks 0 --sample 256 > $OUT/diag2.log 2>&1
tail -3 $OUT/diag2.log
python tools/ncu_summary.py /tmp/f131.ncu-rep > $OUT/force_131k_ncu.txt 2>&1
ncu -i /tmp/f131.ncu-rep --page source --csv --print-source cuda 2>/dev/null | gzip > $OUT/force_131k_cuda.csv.gz
ncu -i /tmp/f131.ncu-rep --page source --csv --print-source sass 2>/dev/null | gzip > $OUT/force_131k_sass.csv.gz
ls -la $OUT
