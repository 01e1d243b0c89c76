"""Generate tests/golden/reference_training.npz by executing the REFERENCE'S OWN run_pinn_training body
(pinn_app/software.py:626-1139, with its colpoint_plot 581-605 and ms_error 241-242) on the synthetic world of
tests/golden/training_world.py.

run_pinn_training is the drop-in boundary: what it derives -- the 11 result files with their (inconsistent) key names,
the stage-2 network size / scl / epsil / loss weights from the stage-1 residual and error, the sampling sizes and epoch
counts of stage 2 -- is host logic the B200 driver mirrors.  jax / optax / tfp / matplotlib are absent, so the function is
lifted out of the file with ``ast`` and run with

    jnp -> numpy;  random.PRNGKey / split -> integer key ids;  plt, make_axes_locatable -> MagicMock
    sol_init_MLP, sol_pred_create, mNN_pred_create, gov_eqn, loss_create, data_func_create, adam_optimizer,
    lbfgs_optimizer, predictF -> the recording stand-ins below (no training: tests/golden/training_world.py)

The body of run_pinn_training itself is executed unmodified; np.savez writes its real files into a temporary directory.
Only the file contents and the recorded calls are written to the repo.
"""
import ast
import collections
import os
import sys
import tempfile
import types
from pathlib import Path
from unittest import mock

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from tests.golden import training_world as W  # noqa: E402

REF = "/root/reference/pinn_app/software.py"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "reference_training.npz")
P = collections.namedtuple("P", "stage version")


def main():
    tree = ast.parse(open(REF).read())
    body = [n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name in ("run_pinn_training", "colpoint_plot", "ms_error")]
    assert len(body) == 3
    rec = collections.OrderedDict()
    st = {"stage": 0}

    def sol_init_MLP(key, n_hl, n_unit):
        st["stage"] += 1
        rec[f"init{st['stage']}"] = np.array([n_hl, n_unit], dtype=np.float64)
        return P(st["stage"], 0)

    def sol_pred_create(limit, scl, epsil, act_s=0):
        rec["pred1"] = np.array([scl, epsil, act_s], dtype=np.float64)
        rec["limit"] = np.array([np.asarray(limit[0]), np.asarray(limit[1])], dtype=np.float64)
        return lambda params, z: W.u_field(1, params.version, z)

    def mNN_pred_create(f_u, limit, scl, epsil, act_s=0):
        rec["pred2"] = np.array([float(np.asarray(scl).reshape(-1)[0]), float(np.asarray(epsil).reshape(-1)[0]), act_s], dtype=np.float64)
        return lambda params, z: W.u_field(2, params.version, z)

    def gov_eqn(f_u, z):
        return W.residual_of(f_u(z), z)

    n_df = {"n": 0}

    def data_func_create(N_col, N_bd, boundary, domain):
        n_df["n"] += 1
        stage = n_df["n"]
        rec[f"dataf{stage}"] = np.array(list(np.asarray(N_col)) + [N_bd], dtype=np.float64)
        calls = {"n": 0}
        r = np.linspace(domain["x_min"], domain["x_max"], 111)
        t = np.linspace(domain["y_min"], domain["y_max"], 111)

        def dataf(key, F, R_add, T_add):
            calls["n"] += 1
            n = int(np.sum(np.asarray(N_col))) + 2 * int(N_bd)
            return {"x_col": W.sampled_points(stage, calls["n"], n), "cond_bd": [[None, None], [None, None]]}

        dataf.R, dataf.T = np.meshgrid(r, t)
        return dataf

    def loss_create(pred, lw, loss_ref=1):
        k = 1 if "lw1" not in rec else 2
        rec[f"lw{k}"] = np.asarray(lw, dtype=np.float64)

        def NN_loss(params, data):
            return 0.0, np.array([W.initial_loss(params.stage)] + [0.0] * (W.N_INFO - 1))

        NN_loss.ref = loss_ref
        return NN_loss

    def adam_optimizer(R_add, T_add, lossf, predf, params, dataf, F, epoch, key_adam, lr=1e-3):
        rec[f"adam{params.stage}"] = np.array([epoch, lr, lossf.ref], dtype=np.float64)
        return P(params.stage, params.version + 1), W.loss_rows(params.stage, "adam", epoch)

    def lbfgs_optimizer(lossf, params, data, epoch):
        rec[f"lbfgs{params.stage}"] = np.array([epoch], dtype=np.float64)
        return P(params.stage, params.version + 1), W.loss_rows(params.stage, "lbfgs", int(epoch / 3))

    def predictF(predf, params, z1, z2):
        return W.weight_map(params.version, np.asarray(z1).shape)

    plt = mock.MagicMock()
    plt.subplots.return_value = (mock.MagicMock(), mock.MagicMock())
    rnd = types.SimpleNamespace(PRNGKey=lambda seed: 1, split=lambda key, n=2: [100 * int(key) + i for i in range(n)])
    ns = {"jnp": np, "np": np, "random": rnd, "plt": plt, "make_axes_locatable": mock.MagicMock(), "Path": Path,
          "sol_init_MLP": sol_init_MLP, "sol_pred_create": sol_pred_create, "mNN_pred_create": mNN_pred_create, "gov_eqn": gov_eqn,
          "data_func_create": data_func_create, "loss_create": loss_create, "adam_optimizer": adam_optimizer,
          "lbfgs_optimizer": lbfgs_optimizer, "predictF": predictF, "sys": sys}
    exec(compile(ast.fix_missing_locations(ast.Module(body=body, type_ignores=[])), REF, "exec"), ns)
    out = {}
    with tempfile.TemporaryDirectory() as d:
        ns["run_pinn_training"](**W.KW, output_dir=d)
        files = sorted(os.listdir(d))
        for f in files:
            z = np.load(os.path.join(d, f), allow_pickle=True)
            for k in z.files:
                out[f"file:{f}:{k}"] = np.asarray(z[k], dtype=np.float64)
    for k, v in rec.items():
        out[f"call:{k}"] = v
    np.savez_compressed(OUT, **out)
    print(len(files), "files:", files)
    for k, v in rec.items():
        print("  ", k, v if v.size < 8 else v.shape)
    print("->", OUT, f"({os.path.getsize(OUT) / 1024:.0f} KB)")


if __name__ == "__main__":
    main()
