"""Generate tests/golden/reference_lbfgs_wrapper.json by executing the REFERENCE'S OWN lbfgs_optimizer and
lbfgs_function (pinn_app/software.py:464-514) with tfp.optimizer.lbfgs_minimize replaced by a recorder.

tfp's optimiser itself is third-party arithmetic (restated from Hager & Zhang, checked against scipy); what the
reference's WRAPPER decides is pinned here: the iteration cap int32(epoch / 3), the tolerance, what the objective closure
returns (the UN-normalised loss_info[0] and the gradient of loss / lref), one loss row and one 'Step: NaN' line per
objective evaluation, the ' Total iterations:' line printing the number of evaluations.  Lifted with ``ast``; jnp ->
numpy; ravel_pytree / jit / jax.debug.callback / grad -> trivial stand-ins; the loss function is a recorder.
"""
import ast
import contextlib
import io
import json
import os
import types

import numpy as np

REF = "/root/reference/pinn_app/software.py"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "reference_lbfgs_wrapper.json")


def run(epoch, n_evals):
    tree = ast.parse(open(REF).read())
    body = [n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name in ("lbfgs_optimizer", "lbfgs_function")]
    assert len(body) == 2
    rec = {"epoch": epoch}

    def lossf(params, data):      # returns (normalised loss, loss_info) like loss_fun (sw:318-379)
        s = float(np.sum(params))
        info = np.array([2.0 + s, 1.5 + s, 0.5, 0.7, 0.8, 0.5])
        return info[0] / 4.0, info

    def grad(f, has_aux=False):   # d(normalised loss)/dparams of the recorder above = 1/4 per entry
        return lambda params, data: (np.full_like(params, 0.25), f(params, data)[1])

    def lbfgs_minimize(value_and_gradients_function, initial_position, tolerance, max_iterations):
        rec["tolerance"], rec["max_iterations"] = float(tolerance), int(max_iterations)
        x = np.asarray(initial_position, dtype=np.float64)
        vals = []
        for k in range(n_evals):
            v, g = value_and_gradients_function(x - 0.1 * k)
            vals.append([float(v), float(np.asarray(g)[0])])
        rec["closure_returns"] = vals
        return types.SimpleNamespace(position=x - 0.1 * (n_evals - 1), num_objective_evaluations=n_evals)

    flat = types.SimpleNamespace(ravel_pytree=lambda p: (np.asarray(p, dtype=np.float64), lambda v: np.asarray(v, dtype=np.float64)))
    out = io.StringIO()
    ns = {"jnp": np, "flat_utl": flat, "jit": lambda f: f, "grad": grad,
          "jax": types.SimpleNamespace(debug=types.SimpleNamespace(callback=lambda fn, x: fn(x))),
          "tfp": types.SimpleNamespace(optimizer=types.SimpleNamespace(lbfgs_minimize=lbfgs_minimize))}
    exec(compile(ast.fix_missing_locations(ast.Module(body=body, type_ignores=[])), REF, "exec"), ns)
    with contextlib.redirect_stdout(out):
        params, loss_all = ns["lbfgs_optimizer"](lossf, np.array([0.3, -0.1, 0.2]), None, epoch)
    rec["n_loss_rows"] = len(loss_all)
    rec["loss_rows"] = [[float(v) for v in np.asarray(r)] for r in loss_all]
    rec["stdout"] = out.getvalue().splitlines()
    rec["final_params"] = [float(v) for v in params]
    return rec


if __name__ == "__main__":
    cases = [run(500, 4), run(1500, 2), run(7, 3), run(2, 1)]
    json.dump(cases, open(OUT, "w"), indent=0)
    for c in cases:
        print(c["epoch"], "->", c["max_iterations"], c["tolerance"], c["n_loss_rows"], c["stdout"][-1])
