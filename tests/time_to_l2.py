"""Time-to-L2 (BASELINE.json metric, second half): wall time of an Adam -> L-BFGS schedule on
fixed points until the relative L2 error on the 111-point/111x111 test grid drops below a
threshold.  Workloads: C1 (1D Poisson, u*=x(1-x)) and R0 (polar Laplace, u*=ln r/ln 0.1).
usage: python tools/time_to_l2.py [C1|R0] [adam_steps] [lbfgs_iters] [lw]"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pinn_based_online_pde_calculator_b200 import PinnEngine
from pinn_based_online_pde_calculator_b200.workloads import init_params, make_points, make_workload

_cpu = "--cpu" in sys.argv
sys.argv = [a for a in sys.argv if a != "--cpu"]
name = sys.argv[1] if len(sys.argv) > 1 else "C1"
n_adam = int(sys.argv[2]) if len(sys.argv) > 2 else 2000
n_lbfgs = int(sys.argv[3]) if len(sys.argv) > 3 else 500
wl = make_workload(name)
if len(sys.argv) > 4:
    wl.lw = float(sys.argv[4])
x_col, x_bd, u_bd = make_points(wl)
if name == "C1":
    grid = np.linspace(0, 1, 111, dtype=np.float32)[:, None]
    exact = grid[:, 0] * (1 - grid[:, 0])
else:
    r, t = np.meshgrid(np.linspace(0.1, 1, 111), np.linspace(0, 1, 111))
    grid = np.stack([r.ravel(), t.ravel()], 1).astype(np.float32)
    exact = np.log(grid[:, 0]) / np.log(0.1)
eng = PinnEngine(wl.net, wl.eq, n_bc=len(x_bd))
eng.set_params(init_params(wl.net))
eng.set_points(x_col, x_bd, u_bd)
eng.set_loss(wl.lw, 1.0)
info0 = eng.loss_grad(want_grad=False)[1]
eng.set_loss(wl.lw, float(info0[0]))


def l2():
    u, _, _ = eng.eval(grid)
    return float(np.linalg.norm(u - exact) / np.linalg.norm(exact))


t0 = time.perf_counter()
eng.adam_init()
hist = []
for k in range(0, n_adam, 250):
    eng.adam_steps(250, 1e-3, want_rows=False)
    e = l2()  # synchronises
    hist.append((time.perf_counter() - t0, "adam", k + 250, e))
for k in range(0, n_lbfgs, 50):
    res, rows = eng.lbfgs(50, 1e-10)
    e = l2()
    hist.append((time.perf_counter() - t0, "lbfgs", k + 50, e, res["evaluations"], res["failed"], rows[-1][0] if rows else None))
    if res["failed"] or res["converged"]:
        break
for h in hist:
    print(h)
for thr in (1e-2, 1e-3):
    hit = [h for h in hist if h[3] < thr]
    print(f"time-to-L2<{thr:g}:", f"{hit[0][0]:.3f} s ({hit[0][1]} {hit[0][2]})" if hit else "not reached")

# ---- the same schedule on the CPU oracle (float64), bounded to the Adam part: time to the same thresholds
if _cpu:
    import torch
    from oracle import reference_oracle as O
    from pinn_based_online_pde_calculator_b200.workloads import unflatten
    torch.set_num_threads(os.cpu_count() or 1)
    net = wl.net
    params = [[torch.tensor(W, dtype=torch.float64), torch.tensor(b, dtype=torch.float64)] for W, b in unflatten(net, init_params(net))]
    limit = [torch.tensor(net.lb, dtype=torch.float64), torch.tensor(net.ub, dtype=torch.float64)]
    f_u = O.sol_pred_create(limit, net.scl, net.epsil, act_s=net.act_first, feature_map=net.feature_map)
    residual = None if net.feature_map == "polar" else O.make_gov_eqn_expr(wl.expr, {1: ("x",), 2: ("x", "y")}[net.d_in])
    lossf = O.loss_create(f_u, torch.tensor([wl.lw, 0.0], dtype=torch.float64), 1.0, residual=residual)
    data = dict(x_col=torch.tensor(x_col, dtype=torch.float64), cond_bd=[[torch.tensor(a, dtype=torch.float64) for a in x_bd],
                [torch.tensor(a, dtype=torch.float64)[:, None] for a in u_bd]])
    lossf.ref = float(lossf(params, data)[1][0])
    st = O.AdamState(params)
    g = torch.tensor(grid, dtype=torch.float64)
    t0 = time.perf_counter()
    hits = {}
    for k in range(n_adam):
        params, info, st = O.adam_minimizer(lossf, params, data, 1e-3, st)
        if (k + 1) % 50 == 0:
            e = float(np.linalg.norm(f_u(params, g).numpy()[:, 0] - exact) / np.linalg.norm(exact))
            for thr in (1e-2, 1e-3):
                if e < thr and thr not in hits:
                    hits[thr] = (time.perf_counter() - t0, k + 1)
            if len(hits) == 2:
                break
    print(f"CPU oracle ({os.cpu_count()} cores, float64): time-to-L2 {hits}; {time.perf_counter() - t0:.1f} s for {k + 1} Adam steps")
