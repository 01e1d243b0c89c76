"""Per-step time of the small problems (the reference's own size R0: 5,200 + 200 points, 6x60 polar network; C1: 1k points,
3x20) -- the launch-bound end of the path.  python tools/small_step.py  (needs a B200; PINN_B200_FUSED_TAIL=0/1)"""
import os
import sys

sys.path.insert(0, ".")
from pinn_based_online_pde_calculator_b200 import PinnEngine  # noqa: E402
from pinn_based_online_pde_calculator_b200.workloads import init_params, make_points, make_workload  # noqa: E402

for name in ("R0", "C1"):
    wl = make_workload(name)
    x_col, x_bd, u_bd = make_points(wl)
    eng = PinnEngine(wl.net, wl.eq, n_bc=len(x_bd))
    eng.set_params(init_params(wl.net))
    eng.set_points(x_col, x_bd, u_bd)
    eng.set_loss(wl.lw, 1.0)
    eng.adam_init()
    eng.adam_steps(50, 1e-3, want_rows=False)
    best = 1e9
    for _ in range(5):
        eng.adam_steps(2000, 1e-3, want_rows=False)
        best = min(best, eng.last_ms() / 2000)
    col_ms, bc_ms = eng.time_kernels(20)
    print(f"{name} ({eng.kernel}, fused_tail={os.environ.get('PINN_B200_FUSED_TAIL', '1')}): {best * 1e3:.1f} us/step, "
          f"col kernel {col_ms * 1e3:.1f} us, bc kernel {bc_ms * 1e3:.1f} us, launches/step {eng.launches_per_adam_step()}", flush=True)
    eng.close()
