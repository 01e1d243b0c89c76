"""Randomised parity sweep (needs a B200): random network shapes, equations, activations and point counts through
whatever kernel `auto` selects, loss_info / gradient / u / f against the float64 oracle at the 1e-5 bar, plus a
bit-reproducibility check.  python tools/fuzz_parity.py [n_cases] [seed]"""
import sys
import time

import numpy as np

sys.path.insert(0, ".")
from oracle import reference_oracle as O  # noqa: E402
from tests.helpers import engine_for, make_problem, oracle_loss_grad, rel_err  # noqa: E402

EXPRS = {
    1: ["u_xx + 2*sin(3*x)", "u_xx + u*u_x - x", "u_x - u"],
    2: ["u_xx + u_yy + 2*y*(1-y) + 2*x*(1-x)", "u_xx + u_yy + 9*u - sin(3*x)*sin(2*y)", "u_y + u*u_x - 0.01*u_xx",
        "u_xx + 2*u_xy + 3*u_yy - u*u_y + x", "u*u_xx + u_yy - u_x", "u_x + 2*u_y - u", "(1+x)*u_xx + (2+y)*u_yy - 1"],
    3: ["u_t - 0.1*(u_xx + u_yy)", "u_xx + u_yy + u_tt - 1", "u*u_xx + u_yy - u_t", "u_t + u_x + u_y"],
}
WIDTHS = [16, 20, 32, 40, 50, 64, 65, 100, 120, 128, 129, 200, 256]


def main():
    n_cases = int(sys.argv[1]) if len(sys.argv) > 1 else 40
    rng = np.random.RandomState(int(sys.argv[2]) if len(sys.argv) > 2 else 7)
    worst = {"info": 0.0, "grad": 0.0, "u": 0.0, "f": 0.0}
    bad = 0
    for case in range(n_cases):
        d_in = int(rng.choice([1, 2, 2, 2, 3, 3]))
        width = int(rng.choice(WIDTHS))
        n_hidden = int(rng.randint(1, 7 if width <= 128 else 5))
        expr = str(rng.choice(EXPRS[d_in]))
        n_col = int(rng.choice([1, 7, 31, 33, 64, 257, 1000, 2500, 4800 if width <= 128 else 1500]))
        n_bc = int(rng.randint(1, 5))
        n_bd = int(rng.choice([1, 17, 64, 130]))
        act_first, act_hidden = int(rng.randint(0, 2)), int(rng.randint(0, 2))
        lb = [float(v) for v in rng.uniform(-1.0, 0.0, d_in)]
        ub = [float(v) for v in rng.uniform(0.5, 2.0, d_in)]
        kw = dict(n_hidden=n_hidden, width=width, d_in=d_in, expr=expr, n_col=n_col, n_bd=n_bd, n_bc=n_bc, lb=lb, ub=ub,
                  act_first=act_first, act_hidden=act_hidden, scl=float(rng.choice([1.0, 2.0, 5.0])),
                  epsil=float(rng.choice([1.0, 0.3])), lw=float(rng.choice([1.0, 0.05])), seed=int(rng.randint(1, 10 ** 6)))
        t0 = time.time()
        pb = make_problem(**kw)
        g_ref, info_ref, f_u, residual = oracle_loss_grad(pb, lref=1.3)
        eng = engine_for(pb, lref=1.3)
        g, info = eng.loss_grad()
        g = g.cpu().numpy().copy()
        g2, info2 = eng.loss_grad()
        u, f, _ = eng.eval(pb["x_col"].numpy())
        fu = lambda z: f_u(pb["params"], z)
        u_ref = fu(pb["x_col"]).numpy()[:, 0]
        f_ref = (O.gov_eqn(fu, pb["x_col"]) if residual is None else residual(fu, pb["x_col"])).numpy()[:, 0]
        # loss_info: the total at the parity bar; a single data term mean((u - u_bd)^2) can be a small difference of O(1)
        # numbers (relative error ~ 2 eps_u |u| / |u - u_bd|), so the individual terms get 1e-3
        # u and f: relative to the norm over the points, but not below an rms of 0.1 (a single point can sit on a zero)
        rel_pts = lambda a, b: float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 0.1 * np.sqrt(len(b))))
        e = {"info": float(abs(info[0] / info_ref[0] - 1)), "grad": rel_err(g, g_ref), "u": rel_pts(u, u_ref),
             "f": rel_pts(f, f_ref)}
        terms = float(np.abs(info / info_ref - 1).max())
        det = np.array_equal(g, g2.cpu().numpy()) and np.array_equal(info, info2)
        ok = det and all(v < 1e-5 for v in e.values()) and terms < 1e-3
        e["info"] = max(e["info"], 0.0)
        bad += 0 if ok else 1
        for k in worst:
            worst[k] = max(worst[k], e[k])
        print(f"[{case:3d}] {'ok ' if ok else 'BAD'} {eng.kernel:11s} d={d_in} {n_hidden}x{width} act=({act_first},{act_hidden}) n_col={n_col:5d} "
              f"n_bc={n_bc} '{expr}': loss {e['info']:.1e} (terms {terms:.1e}) grad {e['grad']:.1e} u {e['u']:.1e} f {e['f']:.1e} det={det} "
              f"({time.time() - t0:.1f}s)", flush=True)
        eng.close()
    print(f"cases {n_cases}, failures {bad}, worst: " + ", ".join(f"{k} {v:.2e}" for k, v in worst.items()))
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
