#!/bin/bash
# A/B of the masked kernels' dense FP64 pass: shuffle-reduction rounds per warp (G6_DENSE_ROUNDS = 4 default, 2, 1, 0)
export G6_B200_WAIT_SECONDS=30
for v in "" _r2 _r1 _r0; do
  echo "== libsapporo$v.so"
  for n in 16384 131072; do timeout 20 ./oracle/g6_latency amuse_b200/csrc/libsapporo$v.so $n 150 2>&1 | grep -E "^ni +(4|42|225|2048):" | tr '\n' ';'; echo; done
done 2>&1 | tee gpurun_out/dense_ab.log
G6_B200_LIB=$PWD/amuse_b200/csrc/libsapporo_r0.so timeout 40 python -m pytest tests/test_gpu_parity.py -q -x -p no:cacheprovider -k "golden_predictor or block_step or ragged or massless or update" 2>&1 | tail -2 | tee -a gpurun_out/dense_ab.log
