import sys, traceback
sys.path.insert(0, __import__("os").path.dirname(__import__("os").path.dirname(__import__("os").path.abspath(__file__))))
import numpy as np
from tests.helpers import *
from tests.test_gpu_parity import CASES
from oracle import reference_oracle as O
for name, kw in CASES.items():
    try:
        pb = make_problem(**kw)
        g_ref, info_ref, f_u, residual = oracle_loss_grad(pb, lref=1.7)
        eng = engine_for(pb, lref=1.7)
        g, info = eng.loss_grad()
        g = g.cpu().numpy()
        u, f, _ = eng.eval(pb["x_col"].numpy())
        fu = lambda z: f_u(pb["params"], z)
        u_ref = fu(pb["x_col"]).numpy()[:, 0]
        f_ref = (O.gov_eqn(fu, pb["x_col"]) if residual is None else residual(fu, pb["x_col"])).numpy()[:, 0]
        print(name, "info", np.abs(info/info_ref-1).max(), "grad", rel_err(g, g_ref), "u", rel_err(u,u_ref), "f", rel_err(f,f_ref), flush=True)
        eng.close()
    except Exception as e:
        traceback.print_exc()
        print(name, "FAILED", e, flush=True)
