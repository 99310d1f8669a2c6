#!/bin/bash
# multi-GPU session (charged N x): bench at N ranks, multi-device ABI + torchrun exchange tests, config 4 through the ABI
TAG=${1:-r02}; NG=${2:-8}
OUT=gpurun_out
mkdir -p $OUT
export G6_B200_WAIT_SECONDS=60
nvidia-smi -L | head -8 > $OUT/gpus_$TAG.txt
echo "== bench $NG GPUs"; timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $NG --steps 3 --warmup 3 --no-cpu-baseline --no-extras > $OUT/bench_n${NG}_$TAG.log 2>&1; grep '^{' $OUT/bench_n${NG}_$TAG.log > $OUT/bench_n${NG}_$TAG.json; python - <<PY
import json
d = json.load(open("$OUT/bench_n${NG}_$TAG.json"))
print("value %.4g frac %.4f e2e %.4g parity %s" % (d["value"], d["frac_fp32_peak_nominal"], d["e2e"]["value"], json.dumps(d["parity"])))
PY
echo "== pytest multi"; timeout 240 python -m pytest tests/test_gpu_multidev.py tests/test_gpu_multi.py -q -s --tb=short 2>&1 | grep -v "^W1\|^\[W" | tail -14 | tee $OUT/pytest_multi_$TAG.log
echo "== config 4 through the ABI"; CONFIG4_ENC_ONLY=1 G6_B200_DEVICES=$NG timeout 200 python tools/config4_abi.py 1048576 200 2>&1 | tail -2 | tee $OUT/config4_abi_$TAG.log
echo "== ABI sweep timing"; G6_B200_DEVICES=$NG G6_B200_TRACE=1 timeout 60 python tools/abi_sweep_timing.py 1048576 8 2>&1 | tail -3 | tee $OUT/abi_sweep_$TAG.log
