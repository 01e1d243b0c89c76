"""Generate tests/golden/reference_schedule.json by executing the REFERENCE'S OWN Adam loop body
(pinn_app/software.py:396-460, adam_optimizer) with recording stand-ins for everything it calls.

The loop decides WHEN to re-sample the collocation set (every 100 steps), WHEN to refresh the sampling weight F
(predictF every 2000 steps), WHEN to halve the learning rate (the 4000-step mean / std test) and how long the tail loop
runs -- host logic the B200 driver (software.adam_optimizer) has to reproduce event for event.  jax / optax are not
installable here, so the function is lifted out of the file with ``ast`` (as gen_validator_golden.py does) and run with

    optax.adam(learning_rate=lr)   -> an object that only remembers lr                (the arithmetic is pinned elsewhere)
    adam_minimizer(...)            -> returns the next row of a FIXED synthetic loss sequence and records (step, lr, data id)
    dataf / predictF / random.split -> recorders
    jnp                            -> numpy (int32, round, abs, mean, std, array, min: same semantics)

The source of the loop itself is executed unmodified.  Only the recorded events are written to the repo.
"""
import ast
import contextlib
import io
import json
import os
import types

import numpy as np

REF = "/root/reference/pinn_app/software.py"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "reference_schedule.json")


def synthetic_row(t: int, n_info: int = 6):
    """loss_info row of global Adam step t: a decaying loss with a plateau between steps 4000 and 8000 (so that the first
    learning-rate test keeps lr and the second one halves it) and a slow final rise (so that the tail loop runs)."""
    base = 1.0 / (1.0 + 1e-3 * min(t, 4200)) + 0.02 * np.sin(0.37 * t)
    if t >= 12000:
        base += 2e-5 * (t - 12000)
    return np.array([base, 0.6 * base, 0.4 * base] + [base / (k + 2) for k in range(n_info - 3)])


def run(epoch: int):
    tree = ast.parse(open(REF).read())
    fn = [n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name == "adam_optimizer"]
    assert len(fn) == 1
    ev = {"resample_at": [], "predictF_at": [], "lr_runs": [], "data_id_runs": [], "n_steps": 0}
    state = {"t": 0, "data_id": 0}

    class Opt:
        def __init__(self, learning_rate):
            self.lr = learning_rate

        def init(self, params):
            return "opt_state"

    def adam_minimizer(lossf, params, data, opt, opt_state):
        t = state["t"]
        if not ev["lr_runs"] or ev["lr_runs"][-1][1] != opt.lr:
            ev["lr_runs"].append([t, opt.lr])
        if not ev["data_id_runs"] or ev["data_id_runs"][-1][1] != data:
            ev["data_id_runs"].append([t, data])
        state["t"] = t + 1
        return params, synthetic_row(t), opt_state

    def dataf(key, F, R_add, T_add):
        state["data_id"] += 1
        ev["resample_at"].append(state["t"])      # number of Adam steps done when the set is drawn
        return state["data_id"]

    dataf.R, dataf.T = np.zeros((3, 3)), np.zeros((3, 3))

    def predictF(predf, params, R, T):
        ev["predictF_at"].append(state["t"])
        return np.ones((3, 3))

    err = io.StringIO()
    ns = {"optax": types.SimpleNamespace(adam=Opt), "adam_minimizer": adam_minimizer, "predictF": predictF, "np": np,
          "random": types.SimpleNamespace(split=lambda key, n: [key + 1] * n), "sys": types.SimpleNamespace(stderr=err),
          "jnp": types.SimpleNamespace(int32=np.int32, round=np.round, abs=np.abs, mean=np.mean, std=np.std, array=np.array, min=np.min)}
    exec(compile(ast.fix_missing_locations(ast.Module(body=fn, type_ignores=[])), REF, "exec"), ns)
    with contextlib.redirect_stderr(err):
        params, loss_all = ns["adam_optimizer"](None, None, "lossf", "predf", "params", dataf, np.ones((3, 3)), epoch, 0, lr=1e-3)
    ev["n_steps"] = state["t"]
    ev["n_rows"] = len(loss_all)
    ev["loss0_checksum"] = float(np.sum(np.array(loss_all)[:, 0]))
    ev["stderr"] = err.getvalue().splitlines()
    ev["epoch"] = epoch
    return ev


if __name__ == "__main__":
    cases = [run(e) for e in (8100, 12500, 250)]
    json.dump(cases, open(OUT, "w"))
    for c in cases:
        print(c["epoch"], "steps", c["n_steps"], "rows", c["n_rows"], "resamples", len(c["resample_at"]), "predictF", c["predictF_at"],
              "lr", c["lr_runs"], "stderr lines", len(c["stderr"]))
