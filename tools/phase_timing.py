"""Per-phase clock64 breakdown of the tensor-core collocation kernel (CTA 0)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pinn_based_online_pde_calculator_b200 import PinnEngine
from pinn_based_online_pde_calculator_b200.workloads import init_params, make_points, make_workload
name = sys.argv[1] if len(sys.argv) > 1 else "C2"
wl = make_workload(name, int(sys.argv[2]) if len(sys.argv) > 2 else None)
eng = PinnEngine(wl.net, wl.eq, n_bc=len(wl.n_bd))
eng.set_params(init_params(wl.net))
eng.set_points(*make_points(wl))
eng.set_loss(wl.lw, 1.0)
eng.loss_grad()
p = eng.phase_profile()
tot = sum(p.values())
print(name, eng.kernel, "total clocks", tot)
for k, v in p.items():
    print(f"  {k:16s} {v:12d} {100 * v / max(tot, 1):5.1f}%")
