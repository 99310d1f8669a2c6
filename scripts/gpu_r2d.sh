#!/bin/bash
TAG=${1:-r02f}
OUT=gpurun_out
mkdir -p $OUT
export G6_B200_WAIT_SECONDS=40
echo "== production lib 1M"; timeout 300 python tools/block_stats.py --n 1048576 --k 16,0 --abi-chunks 2 --sample 512 2>&1 | tail -6 | tee $OUT/prod1m_$TAG.log
echo "== production lib 256k"; timeout 300 python tools/block_stats.py --n 262144 --k 16 --abi-chunks 2 --sample 1024 2>&1 | tail -4 | tee $OUT/prod256k_$TAG.log
echo "== tests"; timeout 900 python -m pytest tests -q --tb=short -p xdist -n 1 -m gpu 2>&1 | tail -15 | tee $OUT/tests_$TAG.log
echo "== latency"; for n in 16384 131072; do timeout 120 ./oracle/g6_latency amuse_b200/csrc/libsapporo.so $n 2>&1 | tail -8; done | tee $OUT/latency_$TAG.log
