"""ncu driver for the experimental tcgen05 family: a few train steps of C2 with PINN_B200_KERNEL=umma."""
import os
import sys

os.environ["PINN_B200_KERNEL"] = "umma"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pinn_based_online_pde_calculator_b200 import PinnEngine  # noqa: E402
from pinn_based_online_pde_calculator_b200.workloads import init_params, make_points, make_workload  # noqa: E402

wl = make_workload("C2", int(sys.argv[1]) if len(sys.argv) > 1 else None)
eng = PinnEngine(wl.net, wl.eq, n_bc=len(wl.n_bd))
eng.set_params(init_params(wl.net))
eng.set_points(*make_points(wl))
eng.set_loss(wl.lw, 1.0)
eng.adam_init()
rows = eng.adam_steps(3, 1e-3)
print(eng.kernel, "loss", rows[:, 0], "ms/step", eng.last_ms() / 3)
