"""Generate tests/golden/reference_sampling.npz by executing the REFERENCE'S OWN samplers
(pinn_app/software.py: colloc2D_set 87-136, data_func_create / dataf 521-577) on numpy, with the random draws
replaced by a reproducible table.

What the data dict contains -- LHS interior points, inverse-CDF points in the border ring (F_bd), the boundary groups
joined into the collocation set, residual-adaptive points drawn from F -- is host logic the B200 driver must reproduce.
jax.random / pyDOE streams cannot be reproduced (and are not installable), so the functions are lifted out of the file with
``ast`` and run with

    jnp                 -> numpy  (linspace, meshgrid, where, arange, hstack, vstack, cumsum, interp, floor, int32, ...)
    random.split(k, n)  -> child ids  10 k + i
    random.uniform(k,s) -> numpy RandomState(k).uniform(size=s)          (a table indexed by the key id)
    lhs(2, n)           -> numpy RandomState(1000 + call index).uniform(size=(n, 2))
    X.at[i, j].set(v)   -> copy-and-assign                                (the one jax-only idiom, rewritten in the AST)

tests/test_host_logic.py feeds the SAME table to software.colloc2D_set / data_func_create (through a fake Key and a
patched lhs) and requires identical arrays.  Only the resulting numbers are written to the repo.
"""
import ast
import os
import types

import numpy as np

REF = "/root/reference/pinn_app/software.py"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "reference_sampling.npz")

# the reference's __main__ smoke arguments (software.py:1143-1190)
DOMAIN = {"x_min": 0.1, "x_max": 1.0, "y_min": 0.0, "y_max": 1.0}
BOUNDARY = {"bd_x1_min": 0.1, "bd_x1_max": 0.1, "bd_y1_min": 0.0, "bd_y1_max": 1.0, "bd_u1": 1.0,
            "bd_x2_min": 1.0, "bd_x2_max": 1.0, "bd_y2_min": 0.0, "bd_y2_max": 1.0, "bd_u2": 0.0}
N_COL, N_BD, ROOT_KEY = [700, 90, 60], 40, 7


def table_uniform(key_id, shape):
    return np.random.RandomState(int(key_id) % (2 ** 31)).uniform(size=tuple(int(s) for s in shape))


class LhsTable:
    def __init__(self):
        self.calls = 0

    def __call__(self, dim, n, *a):
        self.calls += 1
        return np.random.RandomState(1000 + self.calls).uniform(size=(int(n), int(dim)))


class _AtSet(ast.NodeTransformer):
    def __init__(self):
        self.n = 0

    def visit_Call(self, node):
        self.generic_visit(node)
        f = node.func
        if (isinstance(f, ast.Attribute) and f.attr == "set" and isinstance(f.value, ast.Subscript)
                and isinstance(f.value.value, ast.Attribute) and f.value.value.attr == "at"):
            self.n += 1
            return ast.Call(ast.Name("_at_set", ast.Load()), [f.value.value.value, f.value.slice, node.args[0]], [])
        return node


def _at_set(x, idx, v):
    y = np.array(x, copy=True)
    y[idx] = v
    return y


def load():
    tree = ast.parse(open(REF).read())
    body = [n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name in ("colloc2D_set", "data_func_create")]
    assert len(body) == 2
    tr = _AtSet()
    body = [tr.visit(n) for n in body]
    assert tr.n == 1   # F.at[idx[0], idx[1]].set(0)  (sw:532)
    rnd = types.SimpleNamespace(split=lambda key, num=2: [10 * int(key) + i for i in range(num)], uniform=table_uniform)
    lhs = LhsTable()
    ns = {"jnp": np, "jax": types.SimpleNamespace(random=rnd), "random": rnd, "lhs": lhs, "_at_set": _at_set}
    exec(compile(ast.fix_missing_locations(ast.Module(body=body, type_ignores=[])), REF, "exec"), ns)
    return ns, lhs


if __name__ == "__main__":
    ns, lhs = load()
    dataf = ns["data_func_create"](N_COL, N_BD, BOUNDARY, DOMAIN)
    R, T = np.asarray(dataf.R), np.asarray(dataf.T)
    F = 0.5 + np.random.RandomState(99).uniform(size=R.shape) ** 3          # a residual-like sampling weight
    data = dataf(ROOT_KEY, F, R, T)
    # colloc2D_set on its own, on a coarse non-square grid
    x = np.linspace(-1.0, 2.0, 13)
    y = np.linspace(0.0, 1.0, 9)
    X, Y = np.meshgrid(x, y)
    W = np.random.RandomState(5).uniform(size=X.shape)
    pts = ns["colloc2D_set"](31, X, Y, W, 257)
    out = dict(R=R, T=T, F=F, x_col=np.asarray(data["x_col"]), cs_X=X, cs_Y=Y, cs_W=W, cs_pts=np.asarray(pts),
               lhs_calls=lhs.calls)
    for i, (xb, ub) in enumerate(zip(*data["cond_bd"])):
        out[f"x_bd{i}"], out[f"u_bd{i}"] = np.asarray(xb), np.asarray(ub)
    np.savez_compressed(OUT, **out)
    print("x_col", out["x_col"].shape, "bd groups", len(data["cond_bd"][0]), "lhs calls", lhs.calls, "->", OUT,
          f"({os.path.getsize(OUT) / 1024:.0f} KB)")
