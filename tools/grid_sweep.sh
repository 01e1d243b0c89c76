echo "production"; timeout 300 python tools/tc_check.py timing 2>&1 | grep -E "tc:"
for e in 1 2 4 8 15; do echo "TC_EXP=$e"; PINN_B200_LIB=$PWD/pinn_based_online_pde_calculator_b200/libpinn_exp_$e.so timeout 300 python tools/tc_check.py timing 2>&1 | grep -E "tc:"; done
