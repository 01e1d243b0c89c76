import sys; sys.path.insert(0,'/root/repo')
from pinn_based_online_pde_calculator_b200.engine import fma_peak_tflops
for v in (0,4,6,8,9):
    print(v, fma_peak_tflops(0, v))
