"""Generate tests/golden/reference_source_*.npz by EXECUTING THE REFERENCE'S OWN FUNCTION BODIES
(pinn_app/software.py: neural_net 158-184, sol_pred_create 207-218, mNN_pred_create 221-234, ms_error 241-242,
vgmat 246-264, vectgrad 268-279, gov_eqn 283-297, loss_create / loss_fun 310-383, lbfgs_function 464-496 (the value /
gradient closure tfp's L-BFGS calls), predictF 608-623 with gaussian2D_smooth 71-83) on seeded float64 inputs.

jax, optax and tensorflow_probability cannot be installed here, so the reference module cannot be imported.  Its
source text CAN be run: the functions are lifted out of the file with ``ast`` (as gen_validator_golden.py does for
the validator) and executed in a namespace where the ~15 names of the jax API they touch are bound to the float64
torch equivalents (``jnp.tanh -> torch.tanh``, ``vjp -> torch.func.vjp``, ``vmap(f, in_axes=0) -> torch.func.vmap(f,
in_dims=0)``, ``grad(f, has_aux=True) -> torch.func.grad(f, has_aux=True)`` ...).  Two jax-only METHOD idioms have no
torch spelling and are rewritten mechanically in the lifted AST -- nothing else of the source is touched:

    mat = mat.at[l, :, ii].set(1.)        ->   mat = _at_set(mat, (l, slice(None), ii), 1.)      (vgmat, sw:263)
    grad_sol.transpose(1, 0, 2)           ->   grad_sol.permute(1, 0, 2)                          (vectgrad, sw:278)

Nothing of the reference is written into the repo except the numbers it produced.  Run here (the container that has
/root/reference); the tests and the GPU box only read the .npz files.
"""
import ast
import os
import types

import numpy as np
import torch

REF = "/root/reference/pinn_app/software.py"
OUT_DIR = os.path.dirname(os.path.abspath(__file__))
WANT = ["neural_net", "sol_pred_create", "mNN_pred_create", "ms_error", "vgmat", "vectgrad", "gov_eqn", "loss_create",
        "lbfgs_function", "predictF"]          # + gaussian2D_smooth (sw:71-83), executed on numpy / scipy (see below)


class _Idioms(ast.NodeTransformer):
    """the two jax-only method idioms (see the module docstring); counts what it rewrote"""

    def __init__(self):
        self.n_at, self.n_tr = 0, 0

    def visit_Call(self, node):
        self.generic_visit(node)
        f = node.func
        # X.at[IDX].set(V)  ->  _at_set(X, IDX, V)
        if (isinstance(f, ast.Attribute) and f.attr == "set" and isinstance(f.value, ast.Subscript)
                and isinstance(f.value.value, ast.Attribute) and f.value.value.attr == "at"):
            self.n_at += 1
            idx = f.value.slice
            elts = idx.elts if isinstance(idx, ast.Tuple) else [idx]
            conv = [ast.Call(ast.Name("slice", ast.Load()), [ast.Constant(None)], []) if isinstance(e, ast.Slice) else e for e in elts]
            return ast.Call(ast.Name("_at_set", ast.Load()), [f.value.value.value, ast.Tuple(conv, ast.Load()), node.args[0]], [])
        # X.transpose(a, b, c)  ->  X.permute(a, b, c)      (jax: axis permutation; torch.transpose swaps two dims)
        if isinstance(f, ast.Attribute) and f.attr == "transpose" and len(node.args) == 3:
            self.n_tr += 1
            return ast.Call(ast.Attribute(f.value, "permute", ast.Load()), node.args, [])
        return node


def _at_set(x, idx, v):
    y = x.clone()
    y[idx] = v
    return y


def load_reference_functions():
    tree = ast.parse(open(REF).read())
    body = [n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name in WANT]
    assert sorted(n.name for n in body) == sorted(WANT), [n.name for n in body]
    tr = _Idioms()
    body = [tr.visit(n) for n in body]
    assert (tr.n_at, tr.n_tr) == (1, 1), (tr.n_at, tr.n_tr)   # exactly the two documented rewrites
    f64 = torch.float64
    jnp = types.SimpleNamespace(
        tanh=torch.tanh, sin=torch.sin, cos=torch.cos, cosh=torch.cosh, dot=torch.matmul, square=torch.square,
        concatenate=lambda xs, axis=0: torch.cat(list(xs), dim=axis),
        hstack=lambda xs: torch.hstack([x if isinstance(x, torch.Tensor) else torch.as_tensor(x, dtype=f64) for x in xs]),
        mean=lambda x, axis=None: torch.mean(x) if axis is None else torch.mean(x, dim=axis),
        sum=lambda x, axis=None: torch.sum(x) if axis is None else torch.sum(x, dim=axis),
        zeros=lambda shape: torch.zeros(tuple(shape), dtype=f64),
        array=lambda v: torch.stack(list(v)) if (len(v) and isinstance(v[0], torch.Tensor)) else torch.tensor(v, dtype=f64),
        split=lambda z, n, axis=0: torch.split(z, z.shape[axis] // n, dim=axis),
    )
    ns = {"jnp": jnp, "vjp": torch.func.vjp, "vmap": lambda f, in_axes=0: torch.func.vmap(f, in_dims=in_axes),
          "grad": lambda f, has_aux=False: torch.func.grad(f, has_aux=has_aux), "_at_set": _at_set, "slice": slice, "range": range,
          "len": len, "zip": zip}
    # lbfgs_function: ravel_pytree, @jit and jax.debug.callback (sw:466-488)
    def ravel_pytree(tree_):
        leaves = [t for pair in tree_ for t in pair]
        shapes = [t.shape for t in leaves]
        flat = torch.cat([t.reshape(-1) for t in leaves])

        def unflat(v):
            out, o = [], 0
            for i in range(0, len(shapes), 2):
                pair = []
                for sh in shapes[i:i + 2]:
                    n = int(np.prod(sh))
                    pair.append(v[o:o + n].reshape(sh))
                    o += n
                out.append(pair)
            return out

        return flat, unflat

    ns["flat_utl"] = types.SimpleNamespace(ravel_pytree=ravel_pytree)
    ns["jit"] = lambda f: f
    ns["jax"] = types.SimpleNamespace(debug=types.SimpleNamespace(callback=lambda fn, x: fn(x.detach())))
    ns["print"] = lambda *a, **k: None
    # gaussian2D_smooth (sw:71-83) only builds a window and calls scipy.signal.convolve2d: run it on numpy / scipy.stats
    import scipy
    import scipy.signal
    import scipy.stats
    g2 = [n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name == "gaussian2D_smooth"]
    ns_np = {"jnp": types.SimpleNamespace(int32=np.int32, linspace=np.linspace, sum=np.sum), "scipy": scipy,
             "jsp": types.SimpleNamespace(stats=scipy.stats)}
    exec(compile(ast.fix_missing_locations(ast.Module(body=g2, type_ignores=[])), REF, "exec"), ns_np)
    ns["gaussian2D_smooth"] = lambda F, sig, wid: ns_np["gaussian2D_smooth"](np.asarray(F.detach() if isinstance(F, torch.Tensor) else F), sig, wid)
    jnp.reshape = lambda x, shape: torch.reshape(x, tuple(shape))
    mod = ast.Module(body=body, type_ignores=[])
    exec(compile(ast.fix_missing_locations(mod), REF, "exec"), ns)
    return ns


def make_case(seed, n_hl, n_unit, n_col, n_bd, scl, epsil, act_s, lw0, lref):
    g = torch.Generator().manual_seed(seed)
    f32 = lambda t: t.float().double()   # inputs exactly representable in fp32: the CUDA engine sees the same numbers
    layers = [3] + n_hl * [n_unit] + [1]
    params = []
    for i, o in zip(layers[:-1], layers[1:]):
        std = (2.0 / (i + o)) ** 0.5
        params.append([f32(torch.randn(i, o, generator=g, dtype=torch.float64).clamp(-2, 2) * std),
                       f32(torch.randn(o, generator=g, dtype=torch.float64).clamp(-2, 2) * std)])
    lb, ub = torch.tensor([0.1, 0.0], dtype=torch.float64), torch.tensor([1.0, 1.0], dtype=torch.float64)
    x_col = f32(torch.rand(n_col, 2, generator=g, dtype=torch.float64) * (ub - lb) + lb)
    x_bd, u_bd = [], []
    for side, val in ((0.1, 1.0), (1.0, 0.0)):   # the reference's smoke problem: u = 1 at r = 0.1, u = 0 at r = 1
        p = torch.rand(n_bd, 2, generator=g, dtype=torch.float64) * (ub - lb) + lb
        p[:, 0] = side
        x_bd.append(f32(p))
        u_bd.append(torch.full((n_bd, 1), val, dtype=torch.float64))
    return params, [lb, ub], x_col, x_bd, u_bd, dict(scl=scl, epsil=epsil, act_s=act_s, lw0=lw0, lref=lref)


def run_case(ns, name, **kw):
    params, limit, x_col, x_bd, u_bd, c = make_case(**kw)
    f_u = ns["sol_pred_create"](limit, c["scl"], c["epsil"], c["act_s"])          # sw:207-218
    fz = lambda z: f_u(params, z)
    u = fz(x_col)
    ug, _ = ns["vectgrad"](fz, x_col)                                               # sw:268-279
    f = ns["gov_eqn"](fz, x_col)                                                    # sw:283-297
    lossf = ns["loss_create"](f_u, torch.tensor([c["lw0"], 0.0], dtype=torch.float64), c["lref"])   # sw:310-383
    data = {"x_col": x_col, "cond_bd": [x_bd, u_bd]}
    loss_n, loss_info = lossf(params, data)
    grads, info2 = torch.func.grad(lossf, has_aux=True)(params, data)               # sw:390
    assert torch.equal(loss_info, info2)
    # stage 2 (sw:221-234): a second network on top of the frozen first one
    g = torch.Generator().manual_seed(kw["seed"] + 1)
    params2 = [[W + 0.05 * torch.randn(W.shape, generator=g, dtype=torch.float64).float().double(), b] for W, b in params]
    f_comb = ns["mNN_pred_create"](fz, limit, 2.0 * c["scl"], 0.1 * c["epsil"], 1)
    u2 = f_comb(params2, x_col)
    f2 = ns["gov_eqn"](lambda z: f_comb(params2, z), x_col)
    # the closure tfp.optimizer.lbfgs_minimize evaluates (sw:464-496): UN-normalised value, gradient of loss / lref
    flat = torch.cat([t.reshape(-1) for pair in params for t in pair])
    fl = ns["lbfgs_function"](lossf, params, data)
    lb_value, lb_grad = fl(flat)
    # predictF (sw:608-623): residual-based sampling weight on a grid, smoothed by the 5x5 Gaussian window
    r = torch.linspace(0.1, 1.0, 23, dtype=torch.float64)
    t_ = torch.linspace(0.0, 1.0, 19, dtype=torch.float64)
    Rg, Tg = torch.meshgrid(r, t_, indexing="xy")
    Fs = ns["predictF"](f_u, params, Rg, Tg)
    out = dict(lbfgs_value=float(lb_value), lbfgs_grad=lb_grad.numpy(), grid_r=r.numpy(), grid_t=t_.numpy(), predictF=np.asarray(Fs),
               n_layers=len(params), x_col=x_col.numpy(), u=u.numpy(), u_grad=ug.numpy(), f=f.numpy(), loss_n=float(loss_n),
               loss_info=loss_info.numpy(), u_stage2=u2.numpy(), f_stage2=f2.numpy(), **{k: np.float64(v) for k, v in c.items()})
    for i, ((W, b), (gW, gb), (W2, b2)) in enumerate(zip(params, grads, params2)):
        out[f"W{i}"], out[f"b{i}"], out[f"gW{i}"], out[f"gb{i}"], out[f"W2_{i}"], out[f"b2_{i}"] = (
            W.numpy(), b.numpy(), gW.numpy(), gb.numpy(), W2.numpy(), b2.numpy())
    for i, (xb, ub_) in enumerate(zip(x_bd, u_bd)):
        out[f"x_bd{i}"], out[f"u_bd{i}"] = xb.numpy(), ub_.numpy()
    path = os.path.join(OUT_DIR, f"reference_source_{name}.npz")
    np.savez_compressed(path, **out)
    print(name, "loss_info", loss_info.numpy(), "->", path, f"({os.path.getsize(path) / 1024:.0f} KB)")


if __name__ == "__main__":
    ns = load_reference_functions()
    run_case(ns, "smoke_6x60", seed=11, n_hl=6, n_unit=60, n_col=700, n_bd=60, scl=1.0, epsil=1.0, act_s=0, lw0=0.05, lref=1.0)
    run_case(ns, "sin_3x24", seed=12, n_hl=3, n_unit=24, n_col=300, n_bd=20, scl=2.0, epsil=0.5, act_s=1, lw0=1.0, lref=0.37)
