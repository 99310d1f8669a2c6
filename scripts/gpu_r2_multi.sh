#!/bin/bash
# multi-GPU validation: single-process multi-device ABI, torchrun exchange path, short bench
TAG=${1:-r02m}; NG=${2:-2}
OUT=gpurun_out
mkdir -p $OUT
export G6_B200_WAIT_SECONDS=60
nvidia-smi -L | tee $OUT/gpus_$TAG.txt
echo "== multi-device ABI (one process)"; timeout 600 python tests/multidev_worker.py $NG 20000 2>&1 | tail -12 | tee $OUT/multidev_$TAG.log
echo "== pytest multi"; timeout 900 python -m pytest tests/test_gpu_multidev.py tests/test_gpu_multi.py -q -s --tb=short 2>&1 | tail -25 | tee $OUT/pytest_multi_$TAG.log
echo "== latency, $NG devices in one process"; G6_B200_DEVICES=$NG timeout 100 ./oracle/g6_latency amuse_b200/csrc/libsapporo.so 131072 100 2>&1 | tail -9 | tee $OUT/latency_multidev_$TAG.log
echo "== bench $NG GPUs"; timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $NG --steps 2 --warmup 3 > $OUT/bench_n${NG}_$TAG.log 2>&1; grep -v "^W1\|^\[W" $OUT/bench_n${NG}_$TAG.log | tail -12
