#!/bin/bash
# multi-GPU validation (charged N x, keep short): tests of both multi-GPU forms, bench line at N ranks
TAG=${1:-r02}; NG=${2:-2}
OUT=gpurun_out
mkdir -p $OUT
export G6_B200_WAIT_SECONDS=60
echo "== pytest multi"; timeout 300 python -m pytest tests/test_gpu_multidev.py tests/test_gpu_multi.py -q -s --tb=short 2>&1 | grep -v "^W1\|^\[W" | grep -E "MULTI|passed|failed|Error|assert" | tee $OUT/pytest_multi_n${NG}_$TAG.log
echo "== bench $NG GPUs"; timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $NG --steps 3 --warmup 3 --no-cpu-baseline --no-extras > $OUT/bench_n${NG}_$TAG.log 2>&1; grep '^{' $OUT/bench_n${NG}_$TAG.log > $OUT/bench_n${NG}_$TAG.json; python - <<PY
import json
d = json.load(open("$OUT/bench_n${NG}_$TAG.json"))
print("value %.4g frac %.4f e2e %.4g parity %s" % (d["value"], d["frac_fp32_peak_nominal"], d["e2e"]["value"], d["parity"]["ok"]))
PY
