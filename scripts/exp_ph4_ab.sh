# A/B of ph4 (reference integrator through the g6 ABI) with the current and the previous library build
for r in 1 2; do
for which in new old; do
  if [ $which = old ]; then export LD_LIBRARY_PATH=$PWD/amuse_b200/csrc/old; else unset LD_LIBRARY_PATH; fi
  echo -n "$which: "; python tools/ph4_timing.py 16384 0.125 gpu 2>&1 | tail -1 | cut -c1-110
  echo -n "$which: "; python tools/ph4_timing.py 1024 1 gpu 2>&1 | tail -1 | cut -c1-110
done; done
