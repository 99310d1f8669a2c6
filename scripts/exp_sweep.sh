#!/bin/bash
# tuning sweep over experiment builds (amuse_b200/csrc/exp_*.so)
for lib in amuse_b200/csrc/exp_*.so; do
  G6_B200_LIB=$PWD/$lib timeout 120 python tools/quick_force_bench.py "$@" 2>&1 | grep -v Warning
done
