#!/bin/bash
# Second half of round 1: latency tables, caller timings, ncu launch lists (block-step path and bench).
OUT=gpurun_out
mkdir -p $OUT
{
echo "# python tools/latency_floor.py"; timeout 100 python tools/latency_floor.py
for n in 1024 16384 131072; do echo "# oracle/g6_latency amuse_b200/csrc/libsapporo.so $n 300   (C caller; us per block step)"; timeout 120 oracle/g6_latency amuse_b200/csrc/libsapporo.so $n 300; done
echo "# python tools/ph4_timing.py 1024 1.0 both"; timeout 300 python tools/ph4_timing.py 1024 1.0 both
echo "# python tools/ph4_timing.py 16384 0.125 gpu"; timeout 300 python tools/ph4_timing.py 16384 0.125 gpu
echo "# python tools/hermite_timing.py 1024 1.0 ; 16384 0.125 ; 131072 0.03125"; timeout 200 python tools/hermite_timing.py 1024 1.0; timeout 200 python tools/hermite_timing.py 16384 0.125; timeout 300 python tools/hermite_timing.py 131072 0.03125
echo "# python tools/phigrape_timing.py 131072 1.0 1e-4 3000"; timeout 300 python tools/phigrape_timing.py 131072 1.0 1e-4 3000
echo "# python tools/phigrape_timing.py 16384 0.25 1e-4 ; 1024 1.0 1e-4 (b200 | oracle)"; timeout 300 python tools/phigrape_timing.py 16384 0.25 1e-4; timeout 300 python tools/phigrape_timing.py 1024 1.0 1e-4; timeout 300 python tools/phigrape_timing.py 1024 1.0 1e-4 0 oracle
} > $OUT/latency_b.txt 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $OUT/launches_blockstep.csv oracle/g6_latency amuse_b200/csrc/libsapporo.so 131072 3 > /dev/null 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file $OUT/launches_bench.csv python bench.py --n 262144 --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > $OUT/ncu_list_b.log 2>&1
tail -3 $OUT/latency_b.txt; ls -la $OUT | tail -5
