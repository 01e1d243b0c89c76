export TC_CHECK_KERNELS=tc
timeout 300 python tools/tc_check.py C5s 2>&1 | grep -E "tc|determ|rror|Trace|line" 
for e in 4 1; do
  echo "== PINN_TC_SHARE=$e"
  PINN_TC_SHARE=$e timeout 300 python tools/tc_check.py timing5 2>&1 | grep -E "tc:|phases|rror"
done
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -k "C5 or W256 or literal" 2>&1 | tail -5
