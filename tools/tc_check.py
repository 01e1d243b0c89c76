"""Development check of the tcgen05 family D kernel (needs a B200): evaluation, loss and gradient against the
float64 oracle and the mma.sync kernel on a small C4-shaped problem (ragged last tile), layer-by-layer
gradient errors, then a timing comparison on the real workload."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, ".")
from tests.helpers import engine_for, make_problem, oracle_loss_grad, rel_err  # noqa: E402
from oracle import reference_oracle as O  # noqa: E402

CASES = {
    "C4s": dict(n_hidden=6, width=128, d_in=2, expr="u_xx + u_yy + 9*u - sin(3*x)*sin(2*y)", n_col=1000, n_bd=100, n_bc=4,
                lb=[0.0, 0.0], ub=[1.0, 1.0], act_first=1, act_hidden=1, scl=2.0),
    "C5s": dict(n_hidden=5, width=256, d_in=3, expr="u_t - 0.1*(u_xx + u_yy)", n_col=600, n_bd=100, n_bc=5, lb=[0.0, 0.0, 0.0],
                ub=[1.0, 1.0, 1.0]),
    "W128tanh": dict(n_hidden=3, width=100, d_in=2, expr="u_xx + u_yy + 2*y*(1-y) + 2*x*(1-x)", n_col=777, n_bd=50, n_bc=4,
                     lb=[0.0, 0.0], ub=[1.0, 1.0]),
}


def check(name):
    pb = make_problem(**CASES[name])
    g_ref, info_ref, f_u, residual = oracle_loss_grad(pb, lref=1.7)
    fu = lambda z: f_u(pb["params"], z)
    u_ref = fu(pb["x_col"]).numpy()[:, 0]
    f_ref = (O.gov_eqn(fu, pb["x_col"]) if residual is None else residual(fu, pb["x_col"])).numpy()[:, 0]
    res = {}
    for kern in ("mma", "tc"):
        os.environ["PINN_B200_KERNEL"] = kern
        eng = engine_for(pb, lref=1.7)
        u, f, j = eng.eval(pb["x_col"].numpy(), want_jets=True)
        print(f"{name} {kern} ({eng.kernel}): eval u {rel_err(u, u_ref):.2e} f {rel_err(f, f_ref):.2e}", flush=True)
        g, info = eng.loss_grad()
        g = g.cpu().numpy()
        print(f"   loss_info max rel {np.abs(info / info_ref - 1).max():.2e}  grad {rel_err(g, g_ref):.2e}", flush=True)
        g2, info2 = eng.loss_grad()
        print("   deterministic:", np.array_equal(g, g2.cpu().numpy()), np.array_equal(info, info2))
        # per-layer gradient error (ravel order: W0, b0, W1, b1, ...)
        o = 0
        lw = pb["net"].layer_widths
        for li, (i, k) in enumerate(zip(lw[:-1], lw[1:])):
            for nm, sz in (("W", i * k), ("b", k)):
                print(f"      layer {li} {nm}: {rel_err(g[o:o + sz], g_ref[o:o + sz]):.2e}", end="")
                o += sz
            print()
        res[kern] = (u, f, j, g, info)
        eng.close()
    print(f"{name}: tc vs mma  u {rel_err(res['tc'][0], res['mma'][0]):.2e} jets {rel_err(res['tc'][2], res['mma'][2]):.2e} "
          f"grad {rel_err(res['tc'][3], res['mma'][3]):.2e}", flush=True)


def timing(n_col=1 << 20, name="C4"):
    from pinn_based_online_pde_calculator_b200 import PinnEngine
    from pinn_based_online_pde_calculator_b200.workloads import init_params, make_points, make_workload

    wl = make_workload(name, n_col=n_col)
    x_col, x_bd, u_bd = make_points(wl)
    fl = wl.flops_per_point()
    for kern in os.environ.get("TC_CHECK_KERNELS", "mma,tc").split(","):
        os.environ["PINN_B200_KERNEL"] = kern
        eng = PinnEngine(wl.net, wl.eq, n_bc=len(x_bd))
        eng.set_params(init_params(wl.net))
        eng.set_points(x_col, x_bd, u_bd)
        eng.set_loss(wl.lw, 1.0)
        eng.adam_init()
        eng.adam_steps(2, 1e-3)
        eng.adam_steps(5, 1e-3)
        ms = eng.last_ms() / 5
        col_ms, bc_ms = eng.time_kernels(3)
        print(f"{name} {n_col} pts {kern}: {ms:.3f} ms/step  col kernel {col_ms:.3f} ms  bc {bc_ms:.3f} ms  -> "
              f"{fl['col'] * n_col / (col_ms * 1e-3) / 1e12:.1f} algorithmic TFLOP/s", flush=True)
        if kern == "tc":
            print("   phases", eng.phase_profile())
        eng.close()


if __name__ == "__main__":
    what = sys.argv[1:] or ["C4s", "W128tanh", "timing"]
    for w in what:
        if w == "timing":
            timing(int(os.environ.get("TC_CHECK_NCOL", 1 << 20)))
        elif w == "timing5":
            timing(1 << 19, "C5")
        else:
            check(w)
