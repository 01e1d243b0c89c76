"""Equation front end: validator parity with the reference (golden verdicts produced by
the reference's own on_equation_change, tests/golden/gen_validator_golden.py), parser, compiler."""
import json
import math
import os

import numpy as np
import pytest

from pinn_based_online_pde_calculator_b200.equation import (EquationError, compile_equation, evaluate_host, parse,
                                                            validate_reference)

GOLD = os.path.join(os.path.dirname(__file__), "golden", "validator_golden.json")


def test_validator_matches_reference_verdicts():
    cases = json.load(open(GOLD))
    assert len(cases) > 4000
    bad = [c["expr"] for c in cases if validate_reference(c["expr"]) != c["invalid"]]
    assert not bad, bad[:10]


def test_validator_examples_from_survey():
    for ok in ["u_xx + 3*u_yy - 5", "u*u_x", "1/(r**2)*u_tt", "(x+y)*(x-y)", "u_x**2", ""]:
        assert validate_reference(ok) is False
    for bad in ["-u", "sin(x)", "1e-3", "((x))", "t", "z", "pi"]:
        assert validate_reference(bad) is True


def test_strict_parser_rejects_what_the_validator_rejects():
    with pytest.raises(EquationError):
        parse("-u_xx", extended=False)
    with pytest.raises(EquationError):
        parse("sin(x)", extended=False)
    parse("u_xx + 3*u_yy - 5", extended=False)


@pytest.mark.parametrize("expr,d_in,jets,jets_lap", [
    ("u_xx + 2", 1, (1, 1, 0), (1, 1, 0)),
    ("u_xx + u_yy + 2*y*(1-y) + 2*x*(1-x)", 2, (2, 2, 0), (2, 0, 2)),
    ("u_t + u*u_x - 0.003183*u_xx", 2, (2, 1, 0), (2, 1, 0)),
    ("u_rr + 1/r*u_r + 1/(r**2)*u_tt", 2, (2, 2, 0), (2, 0, 2)),
    ("u_xx + 2*u_xy + u_yy", 2, (2, 2, 1), (2, 2, 1)),
    ("u_t - 0.1*(u_xx + u_yy)", 3, (3, 2, 0), (3, 0, 2)),
    ("u_x + u_y", 2, (2, 1, 0), (2, 1, 0)),
    ("u*u_xx + u_yy", 2, (2, 2, 0), (2, 2, 0)),            # quasi-linear: coefficient depends on u
    ("(u_xx + u_yy)**2 - 1", 2, (2, 2, 0), (2, 2, 0)),     # not linear at top level
])
def test_jet_structure_selection(expr, d_in, jets, jets_lap):
    ce = compile_equation(expr, d_in=d_in, combine_second=False)
    assert (ce.n1, ce.n2, ce.mix) == jets
    ce = compile_equation(expr, d_in=d_in)
    assert (ce.n1, ce.n2, ce.mix) == jets_lap


def test_combined_second_order_channel_coefficients():
    ce = compile_equation("u_xx + 3*u_yy - 5", d_in=2)
    assert ce.K == 4 and ce.lap_beta[:2] == [1.0, 3.0] and ce.lap_aux == [-1, -1, -1]
    ce = compile_equation("u_rr + 1/r*u_r + 1/(r**2)*u_tt", d_in=2)
    assert ce.lap_beta[0] == 1.0 and ce.lap_aux[1] >= 0          # 1/r**2 is a per-point column
    ce = compile_equation("u_t - 0.1*(u_xx + u_yy)", d_in=3)
    assert ce.K == 5 and ce.lap_beta[:2] == [-0.1, -0.1]
    # residual value with L = sum beta_i u_ii equals the full expression
    rng = np.random.RandomState(1)
    z = rng.rand(32, 2) + 0.5
    full = compile_equation("x*u_xx/2 - (1+y)*u_yy + u_x*u", d_in=2, combine_second=False)
    lap = compile_equation("x*u_xx/2 - (1+y)*u_yy + u_x*u", d_in=2)
    jets = rng.rand(32, 5)
    want = evaluate_host(full, z, jets)
    L = z[:, 0] / 2 * jets[:, 3] - (1 + z[:, 1]) * jets[:, 4]
    got = evaluate_host(lap, z, np.column_stack([jets[:, 0], jets[:, 1], jets[:, 2], L]))
    assert np.allclose(got, want, rtol=1e-12)


def test_unsupported_derivative_sets_raise():
    ce = compile_equation("u_xx + u_yy + u_zz", d_in=3)   # 3-D Laplacian: one combined second-order channel
    assert (ce.n1, ce.n2, ce.mix) == (3, 0, 2) and ce.lap_beta == [1.0, 1.0, 1.0]
    with pytest.raises(EquationError):
        compile_equation("u_xx*u_yy + u_zz", d_in=3)       # three separate second derivatives: K=7, no kernel
    with pytest.raises(EquationError):
        compile_equation("u_xt", d_in=3)
    with pytest.raises(EquationError):
        compile_equation("u_y", d_in=1)


@pytest.mark.parametrize("expr", [
    "u_xx + u_yy + 2*y*(1-y) + 2*x*(1-x)", "u_y + u*u_x - 0.003183*u_xx", "1/(r**2)*u_yy + u_x/r", "u_x**2 - u**3 + x**0",
    "-(u_x)^2 + sin(pi*x)*cos(y) + exp(-u) - 1.5e-1", "sqrt(u*u + 1) + tanh(u_x) + log(x + 2)", "u**-2 + u**0.5 + 2**3",
    "((x + y) * (x - (y - u))) / (1 + x*x)",
])
def test_bytecode_matches_python_eval(expr):
    ce = compile_equation(expr, d_in=2, combine_second=False)
    rng = np.random.RandomState(0)
    n = 64
    z = rng.rand(n, 2) + 0.5
    jets = rng.rand(n, ce.K) + 0.5
    got = evaluate_host(ce, z, jets)
    env = {"x": z[:, 0], "r": z[:, 0], "y": z[:, 1], "t": z[:, 1], "u": jets[:, 0], "pi": math.pi, "sin": np.sin,
           "cos": np.cos, "exp": np.exp, "log": np.log, "tanh": np.tanh, "sqrt": np.sqrt}
    for i in range(ce.n1):
        env["u_" + "xy"[i]] = jets[:, 1 + i]
        env["u_" + "rt"[i]] = jets[:, 1 + i]
    for i in range(ce.n2):
        env["u_" + "xy"[i] * 2] = jets[:, 1 + ce.n1 + i]
    want = eval(expr.replace("^", "**"), {"__builtins__": {}}, env)
    assert np.allclose(got, want, rtol=1e-12)


def test_program_limits_are_enforced():
    with pytest.raises(EquationError):
        compile_equation("+".join(["u*x"] * 80), d_in=2)          # too long
    deep = "u"
    for _ in range(14):
        deep = f"u*(x+{deep})"
    with pytest.raises(EquationError):
        compile_equation(deep, d_in=2)  # stack too deep


def test_point_terms_are_hoisted_out_of_the_step_program():
    ce = compile_equation("u_xx + u_yy + 2*y*(1-y) + 2*x*(1-x)", d_in=2, combine_second=False)
    assert len(ce.ops) == 5 and ce.n_aux == 1 and ce.n_aux_user == 0 and len(ce.aux_ops) > 0
    ce = compile_equation("u_rr + 1/r*u_r + 1/(r**2)*u_tt", d_in=2, combine_second=False)
    assert ce.n_aux == 2 and len(ce.ops) == 9          # the variable coefficients become columns
    ce = compile_equation("u_t + u*u_x - 0.003183*u_xx", d_in=2)
    assert ce.n_aux == 0 and not ce.aux_ops             # nothing to hoist
    ce = compile_equation("u_xx + aux0*u + x*aux1 - 3", d_in=2)
    assert ce.n_aux_user == 2 and ce.n_aux == 3
    raw = compile_equation("u_xx + u_yy + 2*y*(1-y) + 2*x*(1-x)", d_in=2, hoist=False)
    assert len(raw.ops) == 19 and raw.n_aux == 0
