#!/bin/bash
# sweep of experiment builds of the tcgen05 family D on the C4 workload (needs a B200); usage: tools/tc_sweep.sh lib1.so lib2.so ...
P=$PWD/pinn_based_online_pde_calculator_b200
export TC_CHECK_KERNELS=tc
for lib in "$@"; do
  [ -f $P/$lib ] || continue
  echo "== $lib"
  PINN_B200_LIB=$P/$lib timeout 200 python tools/tc_check.py ${TC_SWEEP_WHAT:-timing} 2>&1 | grep -E "tc:|phases|rror"
done
