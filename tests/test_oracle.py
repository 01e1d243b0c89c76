"""Pins for the float64 oracle (oracle/reference_oracle.py).  The reference ships no
tests or golden vectors, so the restatement is pinned by:
OUTPUTS OF THE REFERENCE'S OWN FUNCTION BODIES (software.py:158-383, executed on a torch shim of the jax calls they
make: tests/golden/gen_reference_shim_golden.py), the analytic solution hard-coded at software.py:815, an independent
closed-form jet propagation, finite differences, torch.optim.Adam, and the committed fixtures."""
import glob
import math
import os

import numpy as np
import pytest
import torch

from oracle import reference_oracle as O
from tests.helpers import make_problem, oracle_loss_grad, rel_err

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def _polar_setup(n_hl=3, n_unit=24, n=300, seed=7):
    gen = torch.Generator().manual_seed(seed)
    params = O.sol_init_MLP(gen, n_hl, n_unit)
    limit = [torch.tensor([0.1, 0.0], dtype=torch.float64), torch.tensor([1.0, 1.0], dtype=torch.float64)]
    z = torch.rand(n, 2, generator=gen, dtype=torch.float64) * torch.tensor([0.9, 1.0]) + torch.tensor([0.1, 0.0])
    return params, limit, z


def test_analytic_solution_has_zero_polar_residual():
    # u* = ln r / ln 0.1 (software.py:815) solves u_rr + u_r/r + u_tt/r^2 = 0 (software.py:296)
    z = torch.rand(200, 2, dtype=torch.float64) * torch.tensor([0.9, 1.0]) + torch.tensor([0.1, 0.0])
    f_u = lambda zz: torch.log(zz[:, 0:1]) / math.log(0.1) + 0.0 * zz[:, 1:2]
    f = O.gov_eqn(f_u, z)
    assert f.abs().max() < 1e-10


@pytest.mark.parametrize("act_s", [0, 1])
def test_nested_vjp_residual_matches_closed_form_jets(act_s):
    params, limit, z = _polar_setup()
    f_u = O.sol_pred_create(limit, 1.3, 0.7, act_s=act_s)
    f = O.gov_eqn(lambda zz: f_u(params, zz), z)
    u, ur, ut, urr, utt = O.jet_forward_closed_form(params, z, limit, 1.3, 0.7, act_s=act_s)
    r = z[:, 0:1]
    f2 = urr + ur / r + utt / r ** 2
    assert rel_err(f.numpy(), f2.numpy()) < 1e-12
    assert rel_err(f_u(params, z).numpy(), u.numpy()) < 1e-14


def test_expression_residual_equals_hard_coded_gov_eqn():
    params, limit, z = _polar_setup()
    f_u = O.sol_pred_create(limit, 1.0, 1.0)
    fu = lambda zz: f_u(params, zz)
    res = O.make_gov_eqn_expr("u_rr + 1/r*u_r + 1/(r**2)*u_tt", ("r", "t"))
    assert rel_err(res(fu, z).numpy(), O.gov_eqn(fu, z).numpy()) < 1e-14


def test_loss_info_layout_and_normalisation():
    pb = make_problem(n_hidden=2, width=8, d_in=2, expr="u_xx + u_yy", n_col=50, n_bd=10, n_bc=3, lb=[0, 0], ub=[1, 1],
                      lw=0.3)
    g1, info1, _, _ = oracle_loss_grad(pb, lref=1.0)
    g2, info2, _, _ = oracle_loss_grad(pb, lref=4.0)
    assert info1.shape == (3 + 3 + 1,)
    assert np.allclose(info1, info2)                       # loss_info is un-normalised (software.py:377)
    assert np.allclose(g1, 4.0 * g2)                       # the gradient is of loss/lref (software.py:375)
    assert np.isclose(info1[1], info1[3:6].sum())          # loss_d = sum of data errors
    assert np.isclose(info1[0], info1[1] + 0.3 * info1[2])  # loss = loss_d + lw0*loss_e (software.py:374)
    assert np.isclose(info1[2], info1[6])


def test_gradient_matches_central_finite_differences():
    pb = make_problem(n_hidden=2, width=6, d_in=2, expr="u_xx + u*u_y - x", n_col=40, n_bd=8, n_bc=2, lb=[0, 0],
                      ub=[1, 2], lw=0.7)
    g, info, f_u, residual = oracle_loss_grad(pb, lref=2.0)
    flat = O.ravel_params(pb["params"])
    lossf = O.loss_create(f_u, torch.tensor([0.7, 0.0], dtype=torch.float64), 2.0, residual=residual)
    data = dict(x_col=pb["x_col"], cond_bd=[pb["x_bd"], pb["u_bd"]])
    rng = np.random.RandomState(0)
    for _ in range(5):
        d = torch.tensor(rng.randn(flat.numel()))
        h = 1e-5
        lp = lossf(O.unravel_params(flat + h * d, pb["params"]), data)[0]
        lm = lossf(O.unravel_params(flat - h * d, pb["params"]), data)[0]
        fd = float((lp - lm) / (2 * h))
        assert abs(fd - float(torch.tensor(g) @ d)) < 1e-6 * max(1.0, abs(fd))


def test_adam_restatement_matches_torch_adam():
    # optax.adam(lr) == Adam(b1=.9,b2=.999,eps=1e-8) with bias correction; torch.optim.Adam is an
    # independent implementation of the same published rule.
    pb = make_problem(n_hidden=2, width=6, d_in=1, expr="u_xx + 2", n_col=30, n_bd=1, n_bc=2, lb=[0], ub=[1])
    _, _, f_u, residual = oracle_loss_grad(pb)
    lossf = O.loss_create(f_u, torch.tensor([1.0, 0.0], dtype=torch.float64), 3.0, residual=residual)
    data = dict(x_col=pb["x_col"], cond_bd=[pb["x_bd"], pb["u_bd"]])
    params = pb["params"]
    st = O.AdamState(params)
    tp = [[W.clone().requires_grad_(True), b.clone().requires_grad_(True)] for W, b in pb["params"]]
    opt = torch.optim.Adam([t for l in tp for t in l], lr=1e-3, betas=(0.9, 0.999), eps=1e-8)
    for _ in range(5):
        params, _, st = O.adam_minimizer(lossf, params, data, 1e-3, st)
        opt.zero_grad()
        lossf(tp, data)[0].backward()
        opt.step()
    a = O.ravel_params(params).numpy()
    b = torch.cat([t.detach().reshape(-1) for l in tp for t in l]).numpy()
    assert rel_err(a, b) < 1e-9


def test_ravel_order_is_W_then_b_per_layer():
    params = [[torch.arange(6.0).reshape(3, 2), torch.tensor([10.0, 11.0])], [torch.tensor([[20.0], [21.0]]), torch.tensor([30.0])]]
    flat = O.ravel_params(params)
    assert flat.tolist() == [0, 1, 2, 3, 4, 5, 10, 11, 20, 21, 30]
    back = O.unravel_params(flat, params)
    assert all(torch.equal(a, b) for la, lb in zip(params, back) for a, b in zip(la, lb))


def test_init_mlp_statistics():
    gen = torch.Generator().manual_seed(1)
    p = O.init_MLP(gen, [3, 400, 400, 1])
    W = p[1][0]
    std = math.sqrt(2.0 / 800)
    assert W.abs().max() <= 2 * std + 1e-12          # truncated at +-2 sigma (software.py:151)
    assert abs(float(W.std()) / std - 0.88) < 0.02     # std of a +-2-truncated normal = 0.8796
    assert float(p[1][1].abs().max()) > 0              # biases are NOT zero (software.py:152)


def test_lhs_is_stratified_and_colloc_respects_distribution():
    rng = np.random.RandomState(3)
    H = O.lhs_classic(2, 50, rng)
    for j in range(2):
        assert sorted(np.floor(H[:, j] * 50).astype(int).tolist()) == list(range(50))
    x = np.linspace(0, 1, 11)
    X, Y = np.meshgrid(x, x)
    F = np.zeros_like(X)
    F[2:4, 5:8] = 1.0
    pts = O.colloc2D_set(rng.rand(500), rng.rand(2, 500), X, Y, F)
    assert pts[:, 0].min() >= 0.5 and pts[:, 0].max() <= 0.8 + 1e-12
    assert pts[:, 1].min() >= 0.2 and pts[:, 1].max() <= 0.4 + 1e-12


@pytest.mark.parametrize("path", sorted(glob.glob(os.path.join(GOLD, "case_*.npz"))))
def test_oracle_reproduces_committed_golden_vectors(path):
    from tests.golden.gen_golden import CASES, LREF

    name = os.path.basename(path)[len("case_"):-len(".npz")]
    if CASES[name]["width"] > 64:
        pytest.skip("wide golden cases are re-derived only by tests/golden/gen_golden.py (CPU suite time)")
    z = np.load(path)
    pb = make_problem(**CASES[name])
    g, info, _, _ = oracle_loss_grad(pb, lref=LREF)
    assert np.allclose(info, z["loss_info"], rtol=1e-12)
    assert rel_err(g, z["grad"]) < 1e-6  # fixture gradients are stored in fp32
    assert np.array_equal(O.ravel_params(pb["params"]).numpy().astype(np.float32), z["params"])


def _load_reference_source_case(path):
    z = np.load(path)
    t = lambda a: torch.tensor(np.asarray(a), dtype=torch.float64)
    n = int(z["n_layers"])
    params = [[t(z[f"W{i}"]), t(z[f"b{i}"])] for i in range(n)]
    params2 = [[t(z[f"W2_{i}"]), t(z[f"b2_{i}"])] for i in range(n)]
    grads = [[t(z[f"gW{i}"]), t(z[f"gb{i}"])] for i in range(n)]
    x_bd = [t(z[f"x_bd{i}"]) for i in range(2)]
    u_bd = [t(z[f"u_bd{i}"]) for i in range(2)]
    limit = [torch.tensor([0.1, 0.0], dtype=torch.float64), torch.tensor([1.0, 1.0], dtype=torch.float64)]
    return z, params, params2, grads, t(z["x_col"]), x_bd, u_bd, limit


@pytest.mark.parametrize("path", sorted(glob.glob(os.path.join(GOLD, "reference_source_*.npz"))))
def test_oracle_matches_the_reference_source_executed_on_a_jax_shim(path):
    """The numbers in tests/golden/reference_source_*.npz were produced by the REFERENCE'S OWN source text -- neural_net,
    sol_pred_create, mNN_pred_create, ms_error, vgmat, vectgrad, gov_eqn, loss_create (software.py:158-383), lifted with ast
    and executed with the jax names they use bound to float64 torch equivalents (gen_reference_shim_golden.py).  The
    oracle restates those functions; it must reproduce every output to float64 round-off."""
    z, params, params2, grads, x_col, x_bd, u_bd, limit = _load_reference_source_case(path)
    scl, epsil, act_s = float(z["scl"]), float(z["epsil"]), int(z["act_s"])
    f_u = O.sol_pred_create(limit, scl, epsil, act_s=act_s)
    fz = lambda zz: f_u(params, zz)
    tol = dict(rtol=1e-11, atol=1e-12)
    assert np.allclose(fz(x_col).numpy(), z["u"], **tol)
    assert np.allclose(O.vectgrad(fz, x_col)[0].numpy(), z["u_grad"], **tol)
    f_ref = z["f"]
    assert np.allclose(O.gov_eqn(fz, x_col).numpy(), f_ref, rtol=1e-9, atol=1e-9 * np.abs(f_ref).max())
    lossf = O.loss_create(f_u, torch.tensor([float(z["lw0"]), 0.0], dtype=torch.float64), float(z["lref"]))
    data = dict(x_col=x_col, cond_bd=[x_bd, u_bd])
    loss_n, info = lossf(params, data)
    assert np.allclose(info.numpy(), z["loss_info"], rtol=1e-10, atol=0) and abs(float(loss_n) / float(z["loss_n"]) - 1) < 1e-10
    g, _ = O.loss_and_grad(lossf, params, data)
    assert rel_err(O.ravel_params(g).numpy(), O.ravel_params(grads).numpy()) < 1e-10
    # the closure tfp's L-BFGS evaluates (software.py:464-496): UN-normalised value, gradient of loss / lref
    fl = O.lbfgs_function(lossf, params, data)
    v, g1d = fl(O.ravel_params(params))
    assert abs(float(v) / float(z["lbfgs_value"]) - 1) < 1e-10 and rel_err(g1d.numpy(), z["lbfgs_grad"]) < 1e-10
    assert abs(float(z["lbfgs_value"]) - float(z["loss_info"][0])) == 0.0       # the reference returns loss_info[0], not loss_n
    # predictF (software.py:608-623) with the reference's gaussian2D_smooth (software.py:71-83)
    Rg, Tg = torch.meshgrid(torch.tensor(z["grid_r"]), torch.tensor(z["grid_t"]), indexing="xy")
    assert np.allclose(O.predictF(f_u, params, Rg, Tg), z["predictF"], rtol=1e-9, atol=0)
    # stage 2 (software.py:221-234)
    f_comb = O.mNN_pred_create(fz, limit, 2.0 * scl, 0.1 * epsil, act_s=1)
    assert np.allclose(f_comb(params2, x_col).numpy(), z["u_stage2"], **tol)
    f2 = O.gov_eqn(lambda zz: f_comb(params2, zz), x_col).numpy()
    assert np.allclose(f2, z["f_stage2"], rtol=1e-9, atol=1e-9 * np.abs(z["f_stage2"]).max())
