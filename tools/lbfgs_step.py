"""L-BFGS leg on the reference's own problem size (R0): wall time per objective evaluation for the device-resident loop,
the host-driven loop and the round-1 host line search.  python tools/lbfgs_step.py [iters]   (needs a B200)"""
import os
import sys
import time

sys.path.insert(0, ".")
from pinn_based_online_pde_calculator_b200 import PinnEngine  # noqa: E402
from pinn_based_online_pde_calculator_b200.workloads import init_params, make_points, make_workload  # noqa: E402

iters = int(sys.argv[1]) if len(sys.argv) > 1 else 200
modes = os.environ.get("LBFGS_MODES", "device,host,legacy").split(",")
for name in ("R0", "C1"):
    wl = make_workload(name)
    x_col, x_bd, u_bd = make_points(wl)
    for mode in modes:
        os.environ["PINN_B200_LBFGS"] = mode
        eng = PinnEngine(wl.net, wl.eq, n_bc=len(x_bd))
        eng.set_params(init_params(wl.net))
        eng.set_points(x_col, x_bd, u_bd)
        eng.set_loss(wl.lw, 1.0)
        eng.set_loss(wl.lw, float(eng.loss_grad(want_grad=False)[1][0]))
        eng.adam_init()
        eng.adam_steps(300, 1e-3, want_rows=False)
        eng.lbfgs(3, 1e-12)   # graph build outside the clock
        t0 = time.perf_counter()
        res, rows = eng.lbfgs(iters, 1e-12)
        dt = time.perf_counter() - t0
        print(f"{name} {mode}: {res['iterations']} iterations, {res['evaluations']} evaluations, {1e6 * dt / max(1, res['evaluations']):.1f} us/evaluation, "
              f"host syncs {eng.lbfgs_host_syncs()}, final loss {res['final_loss']:.3e}", flush=True)
        eng.close()
