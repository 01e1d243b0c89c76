"""Experimental tcgen05 family: loss/gradient parity against the production kernel + step timing (needs a B200)."""
import ctypes as C
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
from pinn_based_online_pde_calculator_b200 import PinnEngine  # noqa: E402
from pinn_based_online_pde_calculator_b200.workloads import init_params, make_points, make_workload  # noqa: E402


def rel(a, b):
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30))


def engine(kernel, wl, pts):
    if kernel:
        os.environ["PINN_B200_KERNEL"] = kernel
    else:
        os.environ.pop("PINN_B200_KERNEL", None)
    eng = PinnEngine(wl.net, wl.eq, n_bc=len(wl.n_bd))
    eng.set_params(init_params(wl.net) * 1.5)
    eng.set_points(*pts)
    eng.set_loss(wl.lw, 1.0)
    return eng


sizes = [int(a) for a in sys.argv[1:]] or [1, 33, 1000, 100_000]
for n in sizes:
    wl = make_workload("C2", n_col=n)
    if os.environ.get("UMMA_CHECK_SMALL_BC", "1") == "1":
        wl.n_bd = [2] * 4  # keep the boundary term (production kernel in both engines) out of the comparison
    pts = make_points(wl)
    e0 = engine(None, wl, pts)
    g0, i0 = e0.loss_grad()
    e1 = engine("umma", wl, pts)
    g1, i1 = e1.loss_grad()
    g0, g1 = g0.cpu().numpy(), g1.cpu().numpy()
    print(f"n_col={n}: kernel {e1.kernel}: loss_info rel {np.abs(i1 - i0).max() / np.abs(i0).max():.2e}  grad rel {rel(g1, g0):.2e}  "
          f"|g| {np.linalg.norm(g0):.3e}  finite {np.isfinite(g1).all()}", flush=True)
    if rel(g1, g0) > 1e-4:
        # per-parameter-block errors
        o = 0
        lw = wl.net.layer_widths
        for k, (a, b) in enumerate(zip(lw[:-1], lw[1:])):
            for name, sz in (("W", a * b), ("b", b)):
                print(f"   layer {k} {name}: rel {rel(g1[o:o + sz], g0[o:o + sz]):.2e}  |ref| {np.linalg.norm(g0[o:o + sz]):.2e} |got| {np.linalg.norm(g1[o:o + sz]):.2e}")
                o += sz
    e0.close()
    if n >= 100_000:
        e1.adam_init()
        for _ in range(3):
            e1.adam_steps(1, 1e-3, want_rows=False)
        e1.sync()
        t0 = time.perf_counter()
        for _ in range(5):
            e1.adam_steps(1, 1e-3, want_rows=False)
        e1.sync()
        dt = (time.perf_counter() - t0) / 5
        clk = (C.c_longlong * 8)()
        e1.lib.pinn_engine_umma_clocks.argtypes = [C.c_void_p, C.c_void_p]
        e1.lib.pinn_engine_umma_clocks(e1.h, clk)
        tot = sum(clk)
        names = ["fwd epilogue", "fwd mma wait", "output+VM+seeds", "restage (R)", "wgrad issue+wait", "wgrad flush", "act backward (B)", "tail"]
        print(f"  adam step: {dt * 1e3:.3f} ms  ({n / dt / 1e6:.1f} M points/s)")
        print("  CTA 0 clocks: " + ", ".join(f"{nm} {100 * c / max(tot, 1):.1f}%" for nm, c in zip(names, clk)) + f"; total {tot / 1.965e6:.3f} ms")
    e1.close()
