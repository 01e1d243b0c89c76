#!/bin/bash
# build libpinn_engine.so; non-zero exit (and no stale library) on failure
set -e
cd "$(dirname "$0")/../pinn_based_online_pde_calculator_b200/csrc"
if ! make -j"$(nproc)" > /tmp/pinn_build.log 2>&1; then
  grep -E "error" /tmp/pinn_build.log | head -20
  rm -f ../libpinn_engine.so
  echo "BUILD FAILED"
  exit 1
fi
grep -E "warning" /tmp/pinn_build.log | head -5 || true
ls -la ../libpinn_engine.so
