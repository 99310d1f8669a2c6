#!/bin/bash
# round 2: time-boxed diagnostics
TAG=${1:-r02b}
OUT=gpurun_out
mkdir -p $OUT
export G6_B200_WAIT_SECONDS=40
echo "== hermite tests"; timeout 300 python -m pytest tests/test_gpu_hermite.py -q -s --tb=short -p xdist -n 1 2>&1 | tail -30 | tee $OUT/hermite_$TAG.log
echo "== block stats 1M"; G6_B200_LIB=$PWD/amuse_b200/csrc/libsapporo_stats.so G6_B200_TRACE=1 timeout 400 python tools/block_stats.py --n 1048576 --k 16,0 --abi-chunks 3 2>&1 | tail -12 | tee $OUT/stats1m_$TAG.log
echo "== block stats 256k"; G6_B200_LIB=$PWD/amuse_b200/csrc/libsapporo_stats.so timeout 300 python tools/block_stats.py --n 262144 --k 16,64 --abi-chunks 3 2>&1 | tail -12 | tee $OUT/stats256k_$TAG.log
echo "== production lib 1M"; timeout 300 python tools/block_stats.py --n 1048576 --k 16 --abi-chunks 0 --sample 256 2>&1 | tail -4 | tee $OUT/prod1m_$TAG.log
echo "== parity tests"; timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_ph4.py tests/test_phigrape_replay.py -q -s --tb=short -p xdist -n 1 -m gpu 2>&1 | tail -60 | tee $OUT/parity_$TAG.log
