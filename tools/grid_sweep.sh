timeout 300 python tools/tc_check.py C4s 2>&1 | grep -E "tc|grad|eval|info"
for w in gacc; do echo "window $w"; PINN_B200_L2_WINDOW=$w timeout 300 python tools/tc_check.py timing 2>&1 | grep -E "tc:|phases"; done
