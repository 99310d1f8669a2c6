#!/bin/bash
# round 2 diagnostics: kernel durations of block steps by grid shape (K = 32 and K = 0), host phase timers
OUT=gpurun_out; mkdir -p $OUT
export G6_B200_WAIT_SECONDS=60
for k in 32 0; do
  G6_B200_KCLOSE=$k timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file $OUT/launches_lat_k$k.csv ./oracle/g6_latency amuse_b200/csrc/libsapporo.so 16384 30 > /dev/null 2>&1
  echo "== K=$k"; python tools/launch_summary.py $OUT/launches_lat_k$k.csv --grid | head -30
done
G6_B200_TRACE=1 timeout 100 ./oracle/g6_latency amuse_b200/csrc/libsapporo.so 16384 300 2>&1 | tail -4
