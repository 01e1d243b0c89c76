"""Tiny driver for ncu: a few train steps of one workload, nothing else.
usage: python tools/profile_step.py [workload] [n_col] [steps]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from pinn_based_online_pde_calculator_b200 import PinnEngine
from pinn_based_online_pde_calculator_b200.workloads import init_params, make_points, make_workload

name = sys.argv[1] if len(sys.argv) > 1 else "C2"
n_col = int(sys.argv[2]) if len(sys.argv) > 2 else None
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
wl = make_workload(name, n_col)
eng = PinnEngine(wl.net, wl.eq, n_bc=len(wl.n_bd))
eng.set_params(init_params(wl.net))
x_col, x_bd, u_bd = make_points(wl)
eng.set_points(x_col, x_bd, u_bd)
eng.set_loss(wl.lw, 1.0)
eng.adam_init()
rows = eng.adam_steps(steps, 1e-3)
print(name, "loss", rows[:, 0], "ms/step", eng.last_ms() / steps)
