"""Shared builders for the parity tests: the same seeded weights / points are
fed to the float64 oracle and to the CUDA engine."""
import numpy as np
import torch

from oracle import reference_oracle as O
from pinn_based_online_pde_calculator_b200 import NetworkSpec, compile_equation
from pinn_based_online_pde_calculator_b200.equation import REFERENCE_POLAR_LAPLACE

COORD_NAMES = {1: ("x",), 2: ("x", "y"), 3: ("x", "y", "t")}


def make_problem(n_hidden, width, d_in, expr, n_col, n_bd, n_bc, lb, ub, feature_map="affine",
                 act_first=0, act_hidden=0, scl=1.0, epsil=1.0, lw=1.0, seed=1234, coord_names=None,
                 weight_scale=1.0):
    gen = torch.Generator().manual_seed(seed)
    net = NetworkSpec(n_hidden=n_hidden, width=width, lb=list(lb), ub=list(ub), scl=scl, epsil=epsil,
                      act_first=act_first, act_hidden=act_hidden, feature_map=feature_map, d_in=d_in)
    params = O.sol_init_MLP(gen, n_hidden, width, n_feat=net.n_feat)
    if weight_scale != 1.0:
        params = [[W * weight_scale, b * weight_scale] for W, b in params]
    lbt = torch.tensor(lb, dtype=torch.float64)
    ubt = torch.tensor(ub, dtype=torch.float64)
    x_col = torch.rand(n_col, d_in, generator=gen, dtype=torch.float64) * (ubt - lbt) + lbt
    x_bd, u_bd = [], []
    for i in range(n_bc):
        xb = torch.rand(n_bd, d_in, generator=gen, dtype=torch.float64) * (ubt - lbt) + lbt
        xb[:, i % d_in] = lb[i % d_in] if (i // d_in) % 2 == 0 else ub[i % d_in]
        x_bd.append(xb)
        u_bd.append(torch.sin(3.0 * xb.sum(1, keepdim=True)) * 0.5 + 0.1 * i)
    # round inputs to fp32 so both sides see identical numbers
    f32 = lambda t: t.float().double()
    params = [[f32(W), f32(b)] for W, b in params]
    x_col, x_bd, u_bd = f32(x_col), [f32(a) for a in x_bd], [f32(a) for a in u_bd]
    eq = compile_equation(expr, d_in=d_in)
    names = coord_names or COORD_NAMES[d_in]
    return dict(net=net, eq=eq, params=params, x_col=x_col, x_bd=x_bd, u_bd=u_bd, lw=lw, expr=expr,
                limit=[lbt, ubt], names=names)


def oracle_loss_grad(pb, lref=1.0):
    net = pb["net"]
    f_u = O.sol_pred_create(pb["limit"], net.scl, net.epsil, act_s=net.act_first, feature_map=net.feature_map,
                            hidden_act=("tanh", "sin")[net.act_hidden])
    residual = None if (pb["expr"] == REFERENCE_POLAR_LAPLACE and net.feature_map == "polar") else \
        O.make_gov_eqn_expr(pb["expr"], pb["names"])
    lossf = O.loss_create(f_u, torch.tensor([pb["lw"], 0.0], dtype=torch.float64), lref, residual=residual)
    data = dict(x_col=pb["x_col"], cond_bd=[pb["x_bd"], pb["u_bd"]])
    grads, info = O.loss_and_grad(lossf, pb["params"], data)
    return O.ravel_params(grads).numpy(), info.numpy(), f_u, residual


def engine_for(pb, lref=1.0, device=0):
    from pinn_based_online_pde_calculator_b200 import PinnEngine

    eng = PinnEngine(pb["net"], pb["eq"], n_bc=len(pb["x_bd"]), device=device)
    eng.set_params(O.ravel_params(pb["params"]).numpy().astype(np.float32))
    eng.set_points(pb["x_col"].numpy(), [a.numpy() for a in pb["x_bd"]], [a.numpy() for a in pb["u_bd"]])
    eng.set_loss(pb["lw"], lref)
    return eng


def rel_err(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300))


def problem_from_workload(name, n_col, n_bd=64):
    """The LITERAL BASELINE.json workload (workloads.make_workload: network, scl, equation string, source terms,
    boundary data) on a slice of its seeded points, in the dict layout of make_problem."""
    from pinn_based_online_pde_calculator_b200.workloads import init_params, make_points, make_workload, unflatten

    wl = make_workload(name, n_col)
    wl.n_bd = [min(n, n_bd) for n in wl.n_bd]
    x_col, x_bd, u_bd = make_points(wl)
    net = wl.net
    t64 = lambda a: torch.tensor(np.asarray(a, dtype=np.float32), dtype=torch.float32).double()
    params = [[t64(W), t64(b)] for W, b in unflatten(net, init_params(net))]
    names = COORD_NAMES[net.d_in]
    if net.d_in == 2 and "u_t" in wl.expr and "u_y" not in wl.expr:
        names = ("x", "t")
    return dict(net=net, eq=wl.eq, params=params, x_col=t64(x_col), x_bd=[t64(a) for a in x_bd],
                u_bd=[t64(a)[:, None] for a in u_bd], lw=wl.lw, expr=wl.expr,
                limit=[torch.tensor(net.lb, dtype=torch.float64), torch.tensor(net.ub, dtype=torch.float64)], names=names)
