"""CPU float64 ORACLE for the PINN residual-loss + gradient hot path.

*** TEST INFRASTRUCTURE ONLY ***
Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
``--impl reference`` legs may import this module.  The product path
(``pinn_based_online_pde_calculator_b200``) never does, and fails loudly when
its CUDA library is missing.

What it is
----------
A functional restatement of the reference's hot path, ``pinn_app/software.py``
(all ``file:line`` cites below are into ``/root/reference/pinn_app/software.py``),
written with ``torch.func.{vjp,vmap,grad}`` so that the nested reverse-mode
structure of the JAX original (``jax.vjp`` / ``jax.vmap`` / ``jax.grad``) is
kept call for call, in float64 (the reference enables x64 at software.py:18).

PINNED AGAINST THE REFERENCE'S OWN SOURCE TEXT (network, derivatives, residual, loss, gradient, stage 2):
  * jax / optax / tensorflow_probability / pyDOE are not installable here, so the reference MODULE cannot be imported
    and it ships no tests or golden vectors.  Its function bodies can be run, though:
    tests/golden/gen_reference_shim_golden.py lifts neural_net, sol_pred_create, mNN_pred_create, ms_error, vgmat,
    vectgrad, gov_eqn and loss_create (software.py:158-383) out of the file with ``ast`` and executes them with the jax
    names they use bound to float64 torch equivalents (two jax-only method idioms rewritten mechanically, listed in that
    script).  tests/test_oracle.py checks this oracle against those outputs to float64 round-off (u, du/dz, f, loss_info,
    d(loss/lref)/dparams, stage-2 u and f) and tests/test_gpu_parity.py checks the CUDA engine against them at 1e-5.
  * further pins: (i) the analytic solution the reference hard-codes at software.py:815 (u* = ln r / ln 0.1 has zero
    polar-Laplace residual), (ii) an independent closed-form forward-jet propagation (``jet_forward_closed_form``
    below) that agrees with the nested-vjp residual to ~1e-14, (iii) central finite differences of the loss.

PARITY UNPINNED (third-party arithmetic only):
  * jax.random (threefry) streams cannot be reproduced: initial weights and sample points are always passed in as
    explicit arrays.
  * optax.adam is restated from its published update rule (b1=0.9, b2=0.999, eps=1e-8, eps_root=0, bias-corrected) and
    checked against torch.optim.Adam; tfp's L-BFGS / Hager-Zhang line search is restated from the paper and checked
    against scipy's L-BFGS-B (final solutions, not iterates).

Generalisations beyond the reference (needed by BASELINE.json configs C1..C5,
see SURVEY.md section 8d) are opt-in keyword arguments whose defaults reproduce
the reference exactly: ``feature_map='polar'``, ``hidden_act='tanh'``, and a
residual given as an expression string instead of the hard-coded ``gov_eqn``.
"""
from __future__ import annotations

import math
from typing import Callable, Dict, List, Optional, Sequence

import numpy as np
import torch
from torch.func import grad, vjp, vmap

DTYPE = torch.float64


# --------------------------------------------------------------------------
# parameters  (software.py:142-154, 193-203)
# --------------------------------------------------------------------------
def truncated_normal(gen: torch.Generator, shape, lo=-2.0, hi=2.0, dtype=DTYPE):
    """Truncated standard normal by inverse-CDF, the same construction jax.random
    .truncated_normal uses (uniform on [Phi(lo), Phi(hi)] then erfinv); the
    uniform stream is torch's, not threefry -- see 'PARITY UNPINNED'."""
    a = math.erf(lo / math.sqrt(2.0))
    b = math.erf(hi / math.sqrt(2.0))
    u = torch.rand(shape, generator=gen, dtype=torch.float64) * (b - a) + a
    out = math.sqrt(2.0) * torch.erfinv(u)
    return out.clamp_(lo, hi).to(dtype)


def init_MLP(gen: torch.Generator, layer_widths: Sequence[int], dtype=DTYPE):
    """software.py:142-154: per layer W (in,out) and b (out,), BOTH
    truncated-normal(-2,2) * sqrt(2/(in+out)) (biases are not zero)."""
    params = []
    for in_dim, out_dim in zip(layer_widths[:-1], layer_widths[1:]):
        std = math.sqrt(2.0 / (in_dim + out_dim))
        W = truncated_normal(gen, (in_dim, out_dim), dtype=dtype) * std
        b = truncated_normal(gen, (out_dim,), dtype=dtype) * std
        params.append([W, b])
    return params


def sol_init_MLP(gen: torch.Generator, n_hl: int, n_unit: int, n_feat: int = 3, dtype=DTYPE):
    """software.py:193-203: layers = [3] + n_hl*[n_unit] + [1]."""
    return init_MLP(gen, [n_feat] + n_hl * [n_unit] + [1], dtype=dtype)


def ravel_params(params) -> torch.Tensor:
    """jax.flatten_util.ravel_pytree order (software.py:466,481,502):
    W0 row-major, b0, W1, b1, ..."""
    return torch.cat([t.reshape(-1) for layer in params for t in layer])


def unravel_params(flat: torch.Tensor, like):
    out, o = [], 0
    for W, b in like:
        nW, nb = W.numel(), b.numel()
        out.append([flat[o:o + nW].reshape(W.shape), flat[o + nW:o + nW + nb].reshape(b.shape)])
        o += nW + nb
    return out


# --------------------------------------------------------------------------
# network  (software.py:158-184, 207-234)
# --------------------------------------------------------------------------
_ACT = {0: torch.tanh, 1: torch.sin, "tanh": torch.tanh, "sin": torch.sin}


def neural_net(params, z, limit, scl, act_s, feature_map: str = "polar", hidden_act="tanh"):
    """software.py:158-184.

    feature_map='polar' is the reference: [2(z0-lb0)/(ub0-lb0)-1, cos z1, sin z1]
    (only lb[0]/ub[0] are used; theta is not normalised), 172-175.
    feature_map='affine' (extension): 2(z_i-lb_i)/(ub_i-lb_i)-1 for every input.
    Only the first layer is scaled by ``scl`` and uses ``act_s`` (170,178); hidden
    layers are tanh (180-181) unless ``hidden_act`` is overridden (extension).
    """
    lb, ub = limit[0], limit[1]
    actv = _ACT[act_s]
    if feature_map == "polar":
        H_r = 2.0 * (z[:, 0:1] - lb[0]) / (ub[0] - lb[0]) - 1.0
        H_cost = torch.cos(z[:, 1:2])
        H_sint = torch.sin(z[:, 1:2])
        H = torch.cat([H_r, H_cost, H_sint], dim=1)
    elif feature_map == "affine":
        H = 2.0 * (z - lb[None, :]) / (ub[None, :] - lb[None, :]) - 1.0
    else:
        raise ValueError(feature_map)
    first, *hidden, last = params
    H = actv((H @ first[0]) * scl + first[1])
    hact = _ACT[hidden_act]
    for layer in hidden:
        H = hact(H @ layer[0] + layer[1])
    return H @ last[0] + last[1]


def sol_pred_create(limit, scl, epsil, act_s=0, feature_map="polar", hidden_act="tanh"):
    """software.py:207-218: u = epsil * NN(z)."""

    def f_u(params, z):
        return epsil * neural_net(params, z, limit, scl, act_s, feature_map, hidden_act)

    return f_u


def mNN_pred_create(f_u, limit, scl, epsil, act_s=0, feature_map="polar", hidden_act="tanh"):
    """software.py:221-234: u = u_prev(z) + epsil * NN2(z); f_u is frozen."""

    def f_comb(params, z):
        return f_u(z) + epsil * neural_net(params, z, limit, scl, act_s, feature_map, hidden_act)

    return f_comb


# --------------------------------------------------------------------------
# derivative machinery  (software.py:241-279)
# --------------------------------------------------------------------------
def ms_error(diff):
    """software.py:241-242."""
    return torch.mean(torch.square(diff), dim=0)


def vgmat(z, n_out, idx=None):
    """software.py:246-264: one-hot cotangents [n_idx, N, n_out]."""
    if idx is None:
        idx = range(n_out)
    idx = list(idx)
    mat = torch.zeros((len(idx), z.shape[0], n_out), dtype=z.dtype)
    for l, ii in enumerate(idx):
        mat[l, :, ii] = 1.0
    return mat


def vectgrad(func, z):
    """software.py:268-279: Jacobian of [N,n_out] wrt [N,n_in] via vjp + vmap."""
    sol, vjp_fn = vjp(func, z)
    mat = vgmat(z, sol.shape[1])
    grad_sol = vmap(vjp_fn, in_dims=0)(mat)[0]
    n_pd = z.shape[1] * sol.shape[1]
    grad_all = grad_sol.permute(1, 0, 2).reshape(z.shape[0], n_pd)
    return grad_all, sol


# --------------------------------------------------------------------------
# residuals
# --------------------------------------------------------------------------
def gov_eqn(f_u, z):
    """software.py:283-297 (the hard-coded polar Laplacian)."""
    u_g, u = vectgrad(f_u, z)
    u_r = u_g[:, 0:1]
    fu_r = lambda zz: vectgrad(f_u, zz)[0][:, 0:1]
    fu_t = lambda zz: vectgrad(f_u, zz)[0][:, 1:2]
    u_rr = vectgrad(fu_r, z)[0][:, 0:1]
    u_tt = vectgrad(fu_t, z)[0][:, 1:2]
    r, t = z[:, 0:1], z[:, 1:2]
    return u_rr + 1 / r * u_r + 1 / (r ** 2) * u_tt


_FUNCS = {"sin": torch.sin, "cos": torch.cos, "exp": torch.exp, "log": torch.log,
          "tanh": torch.tanh, "sqrt": torch.sqrt, "pi": math.pi}


def make_gov_eqn_expr(expr: str, coord_names: Sequence[str], aux: Optional[Callable] = None):
    """Generalised residual: evaluate the user's expression (grammar of
    callbacks/input_validation.py:29-46 plus documented extensions) with
    u, u_a, u_ab obtained by the SAME nested ``vectgrad`` construction gov_eqn
    uses.  ``coord_names[i]`` names input column i (e.g. ('x','y'), ('r','t'),
    ('x','y','t')).  Independent of the product's expression compiler: the
    string is handed to Python ``eval`` over torch tensors.
    """
    import re

    names = set(re.findall(r"u_[a-z]{1,2}", expr))
    code = compile(expr.replace("^", "**"), "<pde>", "eval")
    pos = {n: i for i, n in enumerate(coord_names)}

    def residual(f_u, z, aux_vals: Optional[Dict[str, torch.Tensor]] = None):
        env = dict(_FUNCS)
        for n, i in pos.items():
            env[n] = z[:, i:i + 1]
        u_g, u = vectgrad(f_u, z)
        env["u"] = u
        firsts = {n[2] for n in names if len(n) == 3} | {c for n in names if len(n) == 4 for c in n[2:]}
        for a in firsts:
            env["u_" + a] = u_g[:, pos[a]:pos[a] + 1]
        for n in names:
            if len(n) == 4:
                a, b = n[2], n[3]
                fa = lambda zz, a=a: vectgrad(f_u, zz)[0][:, pos[a]:pos[a] + 1]
                env[n] = vectgrad(fa, z)[0][:, pos[b]:pos[b] + 1]
        if aux_vals:
            env.update(aux_vals)
        out = eval(code, {"__builtins__": {}}, env)
        if not torch.is_tensor(out):
            out = torch.full_like(u, float(out))
        return out + 0 * u

    return residual


# --------------------------------------------------------------------------
# loss  (software.py:310-383)
# --------------------------------------------------------------------------
def loss_create(predf_u, lw, loss_ref, residual: Optional[Callable] = None,
                aux_col: Optional[Dict[str, torch.Tensor]] = None):
    """software.py:310-383.  ``residual`` defaults to ``gov_eqn`` (355)."""

    def loss_fun(params, data):
        f_u = lambda z: predf_u(params, z)
        z_bd = data["cond_bd"][0]
        u_bd = data["cond_bd"][1]
        x_col = data["x_col"]
        norm_err = [ms_error(f_u(z_bd[i]) - u_bd[i]) for i in range(len(z_bd))]  # 334-344
        data_err = torch.hstack(norm_err) if norm_err else torch.zeros(0, dtype=x_col.dtype)
        if residual is None:
            f = gov_eqn(f_u, x_col)  # 355
        else:
            f = residual(f_u, x_col, aux_col)
        eqn_err = torch.hstack([ms_error(f)])  # 358-361
        lw_ = loss_fun.lw
        lref = loss_fun.ref
        loss_data = torch.sum(data_err * 1.0)  # 366,370
        loss_eqn = torch.sum(eqn_err * 1.0)  # 367,371
        loss = loss_data + lw_[0] * loss_eqn  # 374 (lw[1] unused)
        loss_n = loss / lref  # 375
        loss_info = torch.hstack([torch.stack([loss, loss_data, loss_eqn]), data_err, eqn_err])  # 377-378
        return loss_n, loss_info

    loss_fun.ref = loss_ref
    loss_fun.lw = lw
    return loss_fun


def loss_and_grad(lossf, params, data):
    """grad(lossf, has_aux=True)(params, data)  (software.py:390, 479)."""
    grads, loss_info = grad(lossf, has_aux=True)(params, data)
    return grads, loss_info


# --------------------------------------------------------------------------
# Adam (optax.adam restated) (software.py:387-393, 398)
# --------------------------------------------------------------------------
class AdamState:
    def __init__(self, params):
        self.count = 0
        self.mu = [[torch.zeros_like(t) for t in layer] for layer in params]
        self.nu = [[torch.zeros_like(t) for t in layer] for layer in params]


def adam_minimizer(lossf, params, data, lr, opt_state: AdamState,
                   b1=0.9, b2=0.999, eps=1e-8):
    """software.py:387-393 with optax.adam(lr) defaults."""
    grads, loss_info = loss_and_grad(lossf, params, data)
    opt_state.count += 1
    t = opt_state.count
    c1 = 1.0 - b1 ** t
    c2 = 1.0 - b2 ** t
    new_params = []
    for li, layer in enumerate(params):
        new_layer = []
        for ti, p in enumerate(layer):
            g = grads[li][ti]
            m = b1 * opt_state.mu[li][ti] + (1 - b1) * g
            v = b2 * opt_state.nu[li][ti] + (1 - b2) * g * g
            opt_state.mu[li][ti], opt_state.nu[li][ti] = m, v
            upd = -lr * (m / c1) / (torch.sqrt(v / c2) + eps)
            new_layer.append(p + upd)
        new_params.append(new_layer)
    return new_params, loss_info, opt_state


def lbfgs_function(lossf, init_params, data):
    """software.py:464-495: flat-vector closure returning (UN-normalised loss,
    gradient of the NORMALISED loss) -- the reference's quirk, kept."""

    def f(params_1d):
        params = unravel_params(params_1d, init_params)
        grads, loss_info = loss_and_grad(lossf, params, data)
        f.loss.append(loss_info.detach().clone())
        return loss_info[0], ravel_params(grads)

    f.loss = []
    f.update = lambda p1d: unravel_params(p1d, init_params)
    return f


# --------------------------------------------------------------------------
# independent cross-check: closed-form forward jets (NOT the reference's method)
# --------------------------------------------------------------------------
def jet_forward_closed_form(params, z, limit, scl, epsil, act_s=0):
    """(u, u_r, u_t, u_rr, u_tt) of the 'polar' network by Taylor-mode forward
    propagation (SURVEY.md section 8a addendum).  Used only to pin the
    nested-vjp restatement above."""
    lb, ub = limit
    a = 2.0 / (ub[0] - lb[0])
    r, t = z[:, 0:1], z[:, 1:2]
    zero = torch.zeros_like(r)
    h = torch.cat([a * (r - lb[0]) - 1.0, torch.cos(t), torch.sin(t)], 1)
    h_r = torch.cat([a + zero, zero, zero], 1)
    h_t = torch.cat([zero, -torch.sin(t), torch.cos(t)], 1)
    h_rr = torch.cat([zero, zero, zero], 1)
    h_tt = torch.cat([zero, -torch.cos(t), -torch.sin(t)], 1)
    first, *hidden, last = params

    def act_jets(A, A_r, A_t, A_rr, A_tt, kind):
        if kind == 0:
            y = torch.tanh(A)
            d1 = 1 - y * y
            d2 = -2 * y * d1
        else:
            y = torch.sin(A)
            d1 = torch.cos(A)
            d2 = -y
        return (y, d1 * A_r, d1 * A_t, d2 * A_r ** 2 + d1 * A_rr, d2 * A_t ** 2 + d1 * A_tt)

    W, b = first
    J = act_jets(h @ W * scl + b, h_r @ W * scl, h_t @ W * scl, h_rr @ W * scl, h_tt @ W * scl, act_s)
    for W, b in hidden:
        J = act_jets(J[0] @ W + b, J[1] @ W, J[2] @ W, J[3] @ W, J[4] @ W, 0)
    W, b = last
    return tuple(epsil * (J[i] @ W + (b if i == 0 else 0)) for i in range(5))


# --------------------------------------------------------------------------
# sampling helpers (adjacent rows) -- numpy restatements
# --------------------------------------------------------------------------
def lhs_classic(n: int, samples: int, rng: np.random.RandomState) -> np.ndarray:
    """pyDOE.lhs(n, samples) with criterion=None (software.py:553,562), restated
    from pyDOE 0.3.8's published ``_lhsclassic``: stratified (i+U)/N per
    dimension, then an independent permutation per dimension."""
    cut = np.linspace(0, 1, samples + 1)
    u = rng.rand(samples, n)
    a, b = cut[:samples], cut[1:samples + 1]
    rd = u * (b - a)[:, None] + a[:, None]
    H = np.zeros_like(rd)
    for j in range(n):
        order = rng.permutation(range(samples))
        H[:, j] = rd[order, j]
    return H


def gaussian2D_smooth(f: np.ndarray, sig, wid) -> np.ndarray:
    """software.py:71-83."""
    import scipy.signal
    import scipy.stats

    xg = np.linspace(-sig[0], sig[0], int(wid[0]))
    yg = np.linspace(-sig[1], sig[1], int(wid[1]))
    window = scipy.stats.norm.pdf(xg) * scipy.stats.norm.pdf(yg)[:, None]
    win_n = window / np.sum(window)
    return scipy.signal.convolve2d(f, win_n, mode="same")


def colloc2D_set(c01: np.ndarray, frac01: np.ndarray, X, Y, F) -> np.ndarray:
    """software.py:87-136 with the two jax.random.uniform draws passed in
    (c01: [Ns], frac01: [2,Ns])."""
    Xc, Yc, Fc = X[0:-1, 0:-1], Y[0:-1, 0:-1], F[0:-1, 0:-1]
    f = Fc.flatten()
    dx = X[0, 1] - X[0, 0]
    dy = Y[1, 0] - Y[0, 0]
    seq = np.arange(f.shape[0] + 1)
    b = np.hstack([0.0, np.cumsum(f)])
    c = c01 * b[-1]
    posi_rd = np.floor(np.interp(c, b, seq))
    idx_out = np.int32(np.floor(posi_rd / Fc.shape[1]))
    idx_in = np.int32(posi_rd % Fc.shape[1])
    Px = Xc[idx_out, idx_in] + frac01[0] * dx
    Py = Yc[idx_out, idx_in] + frac01[1] * dy
    return np.hstack((Px[:, None], Py[:, None]))


def predictF(predf, params, z1: torch.Tensor, z2: torch.Tensor, residual=None) -> np.ndarray:
    """software.py:608-623."""
    fsol = lambda z: predf(params, z)
    z_star = torch.hstack((z1.flatten()[:, None], z2.flatten()[:, None]))
    f0 = gov_eqn(fsol, z_star) if residual is None else residual(fsol, z_star, None)
    f_sq = f0 ** 2
    f_nm = f_sq / torch.mean(f_sq) + 0.5
    F = f_nm.reshape(z1.shape).detach().numpy()
    return gaussian2D_smooth(F, [1, 1], [5, 5])
