"""Experimental tcgen05 family: evaluation parity against the production kernel + timing (needs a B200)."""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
from pinn_based_online_pde_calculator_b200 import PinnEngine  # noqa: E402
from pinn_based_online_pde_calculator_b200.workloads import init_params, make_workload  # noqa: E402

wl = make_workload("C2")
eng = PinnEngine(wl.net, wl.eq, n_bc=len(wl.n_bd))
eng.set_params(init_params(wl.net) * 1.5)
rng = np.random.RandomState(0)


def rel(a, b):
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30))


for n in (1, 31, 32, 33, 1000, 100_000):
    z = rng.uniform(size=(n, 2)).astype(np.float32)
    os.environ.pop("PINN_B200_KERNEL", None)
    u0, f0, j0 = eng.eval(z, want_jets=True)
    os.environ["PINN_B200_KERNEL"] = "umma"
    u1, f1, j1 = eng.eval(z, want_jets=True)
    print(f"n={n}: rel err u {rel(u1, u0):.2e}  f {rel(f1, f0):.2e}  jets {rel(j1, j0):.2e}   max|du| {np.abs(u1 - u0).max():.2e}", flush=True)

n = 1_000_000
zd = torch.rand(n, 2, device="cuda")
for kern in (None, "umma"):
    if kern:
        os.environ["PINN_B200_KERNEL"] = kern
    else:
        os.environ.pop("PINN_B200_KERNEL", None)
    for _ in range(2):
        eng.eval(zd, want_jets=False)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(5):
        eng.eval(zd, want_jets=False)
    eng.sync()
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / 5
    print(f"eval of {n} points, kernel {kern or 'production'}: {dt * 1e3:.3f} ms  ({n / dt / 1e6:.1f} M points/s)")
    if kern:
        import ctypes as C
        clk = (C.c_longlong * 8)()
        eng.lib.pinn_engine_umma_clocks.argtypes = [C.c_void_p, C.c_void_p]
        eng.lib.pinn_engine_umma_clocks(eng.h, clk)
        tot = sum(clk[:3])
        print(f"  CTA 0 clocks: epilogue {clk[0]} ({100 * clk[0] / tot:.1f}%), mma issue+wait {clk[1]} ({100 * clk[1] / tot:.1f}%), output/VM {clk[2]} "
              f"({100 * clk[2] / tot:.1f}%), total {tot} = {tot / 1.965e6:.3f} ms")
