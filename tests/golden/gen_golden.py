"""Generate the numeric golden fixtures tests/golden/case_*.npz from the float64
oracle (oracle/reference_oracle.py): weights, points, loss_info, flat gradient of
loss/lref, and per-point u / residual.  The GPU tests compare the CUDA engine with
these files, so nothing under oracle/ or /root/reference is needed on the GPU box.
Re-run:  python tests/golden/gen_golden.py
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import reference_oracle as O  # noqa: E402
from tests.helpers import make_problem, oracle_loss_grad  # noqa: E402

CASES = {
    "R0_polar_6x60": dict(n_hidden=6, width=60, d_in=2, expr="u_rr + 1/r*u_r + 1/(r**2)*u_tt", n_col=520, n_bd=100,
                          n_bc=2, lb=[0.1, 0.0], ub=[1.0, 1.0], feature_map="polar", lw=0.05, coord_names=("r", "t")),
    "C1_poisson1d_3x20": dict(n_hidden=3, width=20, d_in=1, expr="u_xx + 2", n_col=500, n_bd=1, n_bc=2, lb=[0.0],
                              ub=[1.0]),
    "C2_poisson2d_4x64": dict(n_hidden=4, width=64, d_in=2, expr="u_xx + u_yy + 2*y*(1-y) + 2*x*(1-x)", n_col=777,
                              n_bd=130, n_bc=4, lb=[0.0, 0.0], ub=[1.0, 1.0]),
    "C3_burgers_8x50": dict(n_hidden=8, width=50, d_in=2, expr="u_y + u*u_x - 0.003183*u_xx", n_col=600, n_bd=100,
                            n_bc=3, lb=[-1.0, 0.0], ub=[1.0, 1.0]),
    "C4_helmholtz_sin_6x128": dict(n_hidden=6, width=128, d_in=2, expr="u_xx + u_yy + 9*u - sin(3*x)*sin(2*y)",
                                   n_col=300, n_bd=60, n_bc=4, lb=[0.0, 0.0], ub=[1.0, 1.0], act_first=1,
                                   act_hidden=1, scl=2.0),
    "C5_heat3d_5x256": dict(n_hidden=5, width=256, d_in=3, expr="u_t - 0.1*(u_xx + u_yy)", n_col=200, n_bd=40,
                            n_bc=5, lb=[0.0, 0.0, 0.0], ub=[1.0, 1.0, 1.0]),
    "mixed_2x32": dict(n_hidden=2, width=32, d_in=2, expr="u_xx + 2*u_xy + 3*u_yy - u*u_y + x", n_col=333, n_bd=50,
                       n_bc=1, lb=[0.0, -1.0], ub=[2.0, 1.0]),
}
LREF = 1.7

if __name__ == "__main__":
    out_dir = os.path.join(ROOT, "tests", "golden")
    os.makedirs(out_dir, exist_ok=True)
    for name, kw in CASES.items():
        pb = make_problem(**kw)
        g, info, f_u, residual = oracle_loss_grad(pb, lref=LREF)
        fu = lambda z: f_u(pb["params"], z)
        u = fu(pb["x_col"]).numpy()[:, 0]
        f = (O.gov_eqn(fu, pb["x_col"]) if residual is None else residual(fu, pb["x_col"])).numpy()[:, 0]
        arrs = dict(params=O.ravel_params(pb["params"]).numpy().astype(np.float32),
                    x_col=pb["x_col"].numpy().astype(np.float32), loss_info=info, grad=g.astype(np.float32), u=u, f=f, lref=LREF,
                    n_bc=len(pb["x_bd"]))
        for i, (a, b) in enumerate(zip(pb["x_bd"], pb["u_bd"])):
            arrs[f"x_bd{i}"] = a.numpy().astype(np.float32)
            arrs[f"u_bd{i}"] = b.numpy().astype(np.float32)[:, 0]
        np.savez_compressed(os.path.join(out_dir, f"case_{name}.npz"), **arrs)
        print(name, info[0], np.linalg.norm(g))
