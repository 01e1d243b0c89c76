"""CPU-only tests of the host side: the C-ABI library loads and exports every symbol
include/pinn_engine.h declares (no compute calls), there is NO CPU fallback, the Adam
schedule of software.py:396-460 is reproduced event for event, samplers and sharding."""
import json
import os
import re

import numpy as np
import pytest

import pinn_based_online_pde_calculator_b200 as pkg
from pinn_based_online_pde_calculator_b200 import software as sw
from pinn_based_online_pde_calculator_b200.engine import EXPORTS, load_library, shard_range
from pinn_based_online_pde_calculator_b200.workloads import init_params, make_points, make_workload, unflatten

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_loads_and_exports_every_declared_symbol():
    lib = load_library()
    hdr = open(os.path.join(ROOT, "include", "pinn_engine.h")).read() + open(os.path.join(ROOT, "include", "pinn_engine_debug.h")).read()
    # the boundary header holds no measurement / probe hooks
    assert not re.search(r"pinn_umma_probe|pinn_fma_peak|phase_profile|umma_clocks|lbfgs_trace", open(os.path.join(ROOT, "include", "pinn_engine.h")).read())
    declared = set(re.findall(r"\b(pinn_[a-z_0-9]+)\s*\(", hdr))
    declared -= {"pinn_eval_cb"}
    assert declared, "no declarations parsed"
    assert declared == set(EXPORTS)
    for sym in declared:
        assert hasattr(lib, sym), sym


def test_no_cpu_fallback_without_a_gpu():
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    wl = make_workload("C1")
    with pytest.raises(RuntimeError, match="no CUDA device"):
        pkg.PinnEngine(wl.net, wl.eq, n_bc=2)


def test_missing_library_fails_loudly(tmp_path):
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        load_library(str(tmp_path / "libpinn_engine.so"))


def test_shard_range_partitions_exactly():
    for n in (0, 1, 7, 1000, 1_000_003):
        for w in (1, 2, 3, 8):
            spans = [shard_range(n, r, w) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1


def test_workload_parameter_counts_match_survey_table():
    # P = F*W+W + (L-1)(W^2+W) + W+1
    want = {"R0": 18601, "C1": 901, "C2": 12737, "C3": 18051, "C4": 83073, "C5": 264449}
    for name, p in want.items():
        wl = make_workload(name)
        assert wl.net.n_params == p
        flat = init_params(wl.net)
        assert flat.shape == (p,) and flat.dtype == np.float32
        layers = unflatten(wl.net, flat)
        assert layers[0][0].shape == (wl.net.n_feat, wl.net.width) and layers[-1][0].shape == (wl.net.width, 1)
    fl = make_workload("C2").flops_per_point()
    assert fl["col"] == 373120 and fl["K"] == 5


def test_workload_points_are_seeded_and_in_domain():
    wl = make_workload("C2", n_col=5000)
    a = make_points(wl)
    b = make_points(wl)
    assert np.array_equal(a[0], b[0])
    assert a[0].min() >= 0 and a[0].max() <= 1 and len(a[1]) == 4
    assert not np.array_equal(a[0], make_points(wl, rank=1)[0])


def test_keys_are_deterministic_and_independent():
    k = sw.Key(1234)
    a, b = k.split(2)
    a2, _ = sw.Key(1234).split(2)
    assert a.rng().uniform() == a2.rng().uniform()
    assert a.rng().uniform() != b.rng().uniform()


def test_data_func_create_mirrors_reference_layout():
    d = {"x_min": 0.1, "x_max": 1, "y_min": 0, "y_max": 1}
    bd = {"bd_x1_min": 0.1, "bd_x1_max": 0.1, "bd_y1_min": 0, "bd_y1_max": 1, "bd_u1": 1,
          "bd_x2_min": 1, "bd_x2_max": 1, "bd_y2_min": 0, "bd_y2_max": 1, "bd_u2": 0}
    np.random.seed(1234)
    dataf = sw.data_func_create([300, 100, 50], 100, bd, d)
    data = dataf(sw.Key(1), dataf.R * 0 + 1, dataf.R, dataf.T)
    # n_col + n_bd(ring) + 2*100 BC points + n_add (software.py:562-569)
    assert data["x_col"].shape == (300 + 100 + 200 + 50, 2)
    ring = data["x_col"][300:400]
    wx, wy = 0.9 / 20, 1 / 20
    on_ring = (ring[:, 0] < 0.1 + wx + 0.9 / 110) | (ring[:, 0] > 1 - wx - 0.9 / 110) | (ring[:, 1] < wy + 1 / 110) | (ring[:, 1] > 1 - wy - 1 / 110)
    assert on_ring.all()
    assert np.allclose(data["cond_bd"][0][0][:, 0], 0.1) and np.allclose(data["cond_bd"][1][0], 1.0)
    assert np.allclose(data["cond_bd"][0][1][:, 0], 1.0) and np.allclose(data["cond_bd"][1][1], 0.0)


def test_gaussian_smooth_preserves_constants_in_the_interior():
    F = np.ones((20, 20))
    S = sw.gaussian2D_smooth(F, [1, 1], [5, 5])
    assert np.allclose(S[2:-2, 2:-2], 1.0) and S[0, 0] < 1.0


class _FakeEngine:
    """Records the call pattern adam_optimizer drives (no GPU)."""

    def __init__(self, n_info=6):
        self.n_info, self.calls, self.t = n_info, [], 0

    def adam_init(self):
        self.calls.append(("init",))

    def adam_steps(self, n, lr, want_rows=True):
        self.calls.append(("steps", n, lr))
        rows = np.zeros((n, self.n_info))
        for i in range(n):
            rows[i, 0] = 1.0 / (1 + self.t) + (0.3 if self.t % 2 else 0.0)
            self.t += 1
        return rows


class _FakeModel:
    def __init__(self):
        self.engine = _FakeEngine()
        self.n_set = 0

    def set_data(self, data):
        self.n_set += 1

    def predict(self, z, want_jets=False):
        return np.zeros(len(z), np.float32), np.ones(len(z), np.float32), None


def test_adam_schedule_events(capsys):
    dataf = lambda key, F, R, T: {}
    dataf.R, dataf.T = np.zeros((4, 4)), np.zeros((4, 4))
    m = _FakeModel()
    loss = sw.adam_optimizer(dataf.R, dataf.T, m, dataf, np.ones((4, 4)), 4100, sw.Key(0), lr=1e-3)
    steps = [c for c in m.engine.calls if c[0] == "steps"]
    # chunks end at 100,200,...; at 1999/3999 (predictF); and at the last step
    ends = np.cumsum([c[1] for c in steps]) - 1
    for e in (100, 200, 1900, 1999, 2000, 3999, 4000, 4099):
        assert e in ends
    assert m.n_set == 1 + 40                          # initial data + one resample per 100 steps (software.py:416-422)
    lrs = [c[2] for c in steps]
    assert lrs[0] == 1e-3 and lrs[-1] == 5e-4         # halved once by the 4000-step test (software.py:430-441)
    err = capsys.readouterr().err
    assert "Step: 100 | Loss: " in err and "learning rate for Adam: 5.0000e-04" in err
    assert len(loss) >= 4100


class _FakeAsyncEngine(_FakeEngine):
    """An engine with the non-blocking pair adam_steps_begin / adam_steps_end: records what the host does in between."""

    def adam_steps_begin(self, n, lr):
        self.calls.append(("begin", n, lr))
        self.in_flight = True
        return True

    def adam_steps_end(self, n):
        self.calls.append(("end", n))
        self.in_flight = False
        return _FakeEngine.adam_steps(self, n, 0.0)[:n]


def test_adam_schedule_samples_the_next_set_while_the_steps_run(capsys):
    """Same events, same key sequence and the same number of resamples as the blocking schedule; every resample of the
    Adam loop is computed while the enqueued steps are in flight (software.py:416-422)."""
    seen = {"blocking": [], "async": []}

    def run(engine, tag):
        m = _FakeModel()
        m.engine = engine

        def dataf(key, F, R, T):
            seen[tag].append(((key.ss.entropy, tuple(key.ss.spawn_key)), getattr(engine, "in_flight", False)))
            return {}

        dataf.R, dataf.T = np.zeros((4, 4)), np.zeros((4, 4))
        loss = sw.adam_optimizer(dataf.R, dataf.T, m, dataf, np.ones((4, 4)), 2100, sw.Key(0), lr=1e-3)
        return m, loss

    m0, l0 = run(_FakeEngine(), "blocking")
    m1, l1 = run(_FakeAsyncEngine(), "async")
    capsys.readouterr()
    assert m0.n_set == m1.n_set == 1 + 20
    assert [k for k, _ in seen["blocking"]] == [k for k, _ in seen["async"]]      # identical key sequence
    assert not any(f for _, f in seen["blocking"])
    assert all(f for _, f in seen["async"][1:]) and not seen["async"][0][1]        # all but the initial set: overlapped
    assert np.array_equal(np.array(l0)[:, 0], np.array(l1)[:, 0])
    begins = [c for c in m1.engine.calls if c[0] == "begin"]
    assert sum(c[1] for c in begins) >= 2100


def test_samplers_equal_the_reference_source_on_the_same_random_table(monkeypatch):
    """tests/golden/reference_sampling.npz: the reference's own colloc2D_set (software.py:87-136) and data_func_create /
    dataf (software.py:521-577), lifted with ast and run on numpy with the random draws taken from a reproducible table
    (gen_reference_sampling_golden.py).  The driver's samplers get the SAME table -- a fake Key whose children have the
    ids 10 k + i and draw RandomState(id).uniform, a patched lhs -- and must return identical arrays: the LHS interior
    points, the border-ring points, the boundary groups joined into the collocation set, the residual-adaptive points."""
    from tests.golden.gen_reference_sampling_golden import BOUNDARY, DOMAIN, N_BD, N_COL, ROOT_KEY, LhsTable, table_uniform

    class FakeKey:
        def __init__(self, kid):
            self.kid = kid

        def split(self, num=2):
            return [FakeKey(10 * self.kid + i) for i in range(num)]

        def rng(self):
            kid = self.kid

            class G:
                @staticmethod
                def uniform(size):
                    return table_uniform(kid, (size,) if np.isscalar(size) else size)

            return G

    z = np.load(os.path.join(os.path.dirname(__file__), "golden", "reference_sampling.npz"))
    pts = sw.colloc2D_set(FakeKey(31), z["cs_X"], z["cs_Y"], z["cs_W"], 257)
    assert np.array_equal(pts, z["cs_pts"])
    table = LhsTable()
    monkeypatch.setattr(sw, "lhs", table)
    dataf = sw.data_func_create(N_COL, N_BD, BOUNDARY, DOMAIN, sampler="host")
    assert np.array_equal(dataf.R, z["R"]) and np.array_equal(dataf.T, z["T"])
    data = dataf(FakeKey(ROOT_KEY), z["F"], dataf.R, dataf.T)
    assert table.calls == int(z["lhs_calls"])
    assert data["x_col"].shape == z["x_col"].shape and np.allclose(data["x_col"], z["x_col"], rtol=0, atol=1e-15)
    for i in range(2):
        assert np.allclose(data["cond_bd"][0][i], z[f"x_bd{i}"], rtol=0, atol=1e-15)
        assert np.array_equal(np.asarray(data["cond_bd"][1][i]), z[f"u_bd{i}"])


def test_run_pinn_training_equals_the_reference_body_on_a_synthetic_world(tmp_path, monkeypatch, capsys):
    """tests/golden/reference_training.npz: the REFERENCE'S OWN run_pinn_training body (software.py:626-1139, lifted with ast)
    executed on the stand-ins of tests/golden/training_world.py -- no training, but everything the function DERIVES: the 11
    result files with their key names and contents, the stage-2 network (6x50, sin first layer), its scl / epsil / loss
    weight computed from the stage-1 residual and error, the doubled sampling sizes, the tripled epoch counts, the loss
    reference taken from the first evaluation.  The B200 driver runs on the same stand-ins and must derive the same."""
    import types

    from tests.golden import training_world as W

    gold = np.load(os.path.join(os.path.dirname(__file__), "golden", "reference_training.npz"))
    rec = {}

    class FakeModel:
        def __init__(self, net, eq, n_bc, lw_eqn, device=0, base=None):
            self.stage, self.version, self.ref = (1 if base is None else 2), 0, 1.0
            rec[f"init{self.stage}"] = [net.n_hidden, net.width]
            rec[f"pred{self.stage}"] = [net.scl, net.epsil, net.act_first]
            rec[f"lw{self.stage}"] = lw_eqn
            assert n_bc == 2 and net.feature_map == "polar" and list(net.lb) == [0.1, 0.0] and list(net.ub) == [1.0, 1.0]
            self.engine = types.SimpleNamespace(set_params=lambda p: None)

        def set_data(self, data):
            pass

        def set_ref(self, ref):
            self.ref = float(ref)

        def loss_info(self):
            return np.array([W.initial_loss(self.stage)] + [0.0] * (W.N_INFO - 1))

        def predict(self, z, want_jets=False):
            u = W.u_field(self.stage, self.version, z)
            return u[:, 0], W.residual_of(u, z)[:, 0], None

        def close(self):
            pass

    def fake_adam(Rg, Tg, model, dataf, Fg, epoch, key, lr=1e-3):
        rec[f"adam{model.stage}"] = [epoch, lr, model.ref]
        model.version += 1
        return W.loss_rows(model.stage, "adam", epoch)

    def fake_lbfgs(model, epoch, *a, **k):
        rec[f"lbfgs{model.stage}"] = [epoch]
        model.version += 1
        return W.loss_rows(model.stage, "lbfgs", int(epoch / 3)), None

    n_df = {"n": 0}

    def fake_data_func_create(N_col, N_bd, boundary, domain, **kw):
        n_df["n"] += 1
        stage, calls = n_df["n"], {"n": 0}
        rec[f"dataf{stage}"] = [float(v) for v in np.asarray(N_col)] + [float(N_bd)]

        def dataf(key, F, R_add, T_add):
            calls["n"] += 1
            n = int(np.sum(np.asarray(N_col))) + 2 * int(N_bd)
            return {"x_col": W.sampled_points(stage, calls["n"], n), "cond_bd": [[None, None], [None, None]]}

        dataf.R, dataf.T = np.meshgrid(np.linspace(domain["x_min"], domain["x_max"], 111), np.linspace(domain["y_min"], domain["y_max"], 111))
        return dataf

    monkeypatch.setattr(sw, "Model", FakeModel)
    monkeypatch.setattr(sw, "adam_optimizer", fake_adam)
    monkeypatch.setattr(sw, "lbfgs_optimizer", fake_lbfgs)
    monkeypatch.setattr(sw, "predictF", lambda model, R, T: W.weight_map(model.version, np.asarray(R).shape))
    monkeypatch.setattr(sw, "data_func_create", fake_data_func_create)
    monkeypatch.setattr(sw, "init_params", lambda net, seed=0: np.zeros(1, np.float32))
    sw.run_pinn_training(**W.KW, output_dir=str(tmp_path), sampler="host")
    capsys.readouterr()
    # ---- the 11 files: names, keys, contents
    want = {}
    for k in gold.files:
        if k.startswith("file:"):
            _, fname, key = k.split(":")
            want.setdefault(fname, {})[key] = gold[k]
    assert sorted(os.listdir(tmp_path)) == sorted(want)
    for fname, keys in want.items():
        z = np.load(tmp_path / fname)
        assert sorted(z.files) == sorted(keys), (fname, z.files)
        for key, ref in keys.items():
            got = np.asarray(z[key], dtype=np.float64)
            assert got.shape == ref.shape, (fname, key, got.shape, ref.shape)
            assert np.allclose(got, ref, rtol=1e-9, atol=1e-12), (fname, key, float(np.abs(got - ref).max()))
    # ---- what the orchestration derived and passed on
    call = lambda k: gold[f"call:{k}"]
    for stage in (1, 2):
        assert rec[f"init{stage}"] == list(call(f"init{stage}"))                        # [hidden layers, units] (sw:695, 941-942)
        assert np.allclose(rec[f"pred{stage}"], call(f"pred{stage}"), rtol=1e-9)        # scl, epsil, first-layer activation
        assert np.isclose(rec[f"lw{stage}"], call(f"lw{stage}")[0], rtol=1e-9)          # weight of the equation term
        assert rec[f"dataf{stage}"] == list(call(f"dataf{stage}"))                      # N_col (x2 in stage 2), N_bd (x2)
        assert np.allclose(rec[f"adam{stage}"], call(f"adam{stage}"), rtol=1e-12)       # epochs (x3), lr, NN_loss.ref
        assert rec[f"lbfgs{stage}"] == list(call(f"lbfgs{stage}"))


@pytest.mark.parametrize("case", json.load(open(os.path.join(os.path.dirname(__file__), "golden", "reference_lbfgs_wrapper.json"))),
                         ids=lambda c: f"epoch{c['epoch']}")
def test_lbfgs_wrapper_equals_the_reference_source(case, capsys):
    """tests/golden/reference_lbfgs_wrapper.json: the reference's own lbfgs_optimizer / lbfgs_function (software.py:464-514,
    lifted with ast) around a RECORDING tfp.optimizer.lbfgs_minimize: iteration cap int32(epoch / 3), tolerance 1e-10, the
    closure returning the UN-normalised loss_info[0], one loss row and one 'Step: NaN' line per objective evaluation,
    ' Total iterations:' printing the number of evaluations.  The driver's wrapper must hand the engine the same."""
    got = {}
    rows_ref = np.array(case["loss_rows"])

    class Eng:
        def lbfgs(self, max_iter, tol, value_unnormalised, on_eval):
            got.update(max_iter=max_iter, tol=tol, value_unnormalised=value_unnormalised)
            for r in rows_ref:
                on_eval(r)
            return {"evaluations": len(rows_ref), "iterations": 0, "converged": False, "failed": False, "final_loss": float(rows_ref[-1, 0])}, rows_ref

    class Mdl:
        engine = Eng()

    rows, res = sw.lbfgs_optimizer(Mdl(), case["epoch"])
    out = capsys.readouterr().out.splitlines()
    assert got["max_iter"] == case["max_iterations"] and got["tol"] == case["tolerance"] and got["value_unnormalised"] is True
    assert [v for v, _ in case["closure_returns"]] == [r[0] for r in case["loss_rows"]]     # the reference returns loss_info[0]
    assert len(rows) == case["n_loss_rows"] and out == case["stdout"]


def test_signature_matches_the_reference_call_site():
    """tests/golden/reference_call_site.json: the keyword names the Dash callback passes (pinn_app/callbacks/training.py,
    the run_pinn_training(...) call inside the daemon-thread target) and the reference function's own parameter list
    (software.py:626-638), both read from the reference source with ast.  The drop-in takes exactly these, in this order,
    as its leading parameters; everything it adds is keyword-only with a default."""
    import inspect

    ref = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "reference_call_site.json")))
    assert ref["call_site_keywords"] == ref["parameters"]
    sig = inspect.signature(sw.run_pinn_training)
    lead = [p.name for p in sig.parameters.values() if p.kind is inspect.Parameter.POSITIONAL_OR_KEYWORD]
    assert lead == ref["parameters"]
    extras = [p for p in sig.parameters.values() if p.kind is not inspect.Parameter.POSITIONAL_OR_KEYWORD]
    assert extras and all(p.kind is inspect.Parameter.KEYWORD_ONLY and p.default is not inspect.Parameter.empty for p in extras)


def _synthetic_row(t, n_info=6):   # the loss sequence of tests/golden/gen_reference_schedule_golden.py
    base = 1.0 / (1.0 + 1e-3 * min(t, 4200)) + 0.02 * np.sin(0.37 * t)
    if t >= 12000:
        base += 2e-5 * (t - 12000)
    return np.array([base, 0.6 * base, 0.4 * base] + [base / (k + 2) for k in range(n_info - 3)])


@pytest.mark.parametrize("case", json.load(open(os.path.join(os.path.dirname(__file__), "golden", "reference_schedule.json"))),
                         ids=lambda c: f"epoch{c['epoch']}")
def test_adam_schedule_equals_the_reference_loop_body(case, capsys):
    """tests/golden/reference_schedule.json holds the events of the REFERENCE'S OWN Adam loop (software.py:396-460, lifted
    with ast and run on recording stand-ins, gen_reference_schedule_golden.py) for a fixed synthetic loss sequence: when it
    re-samples, when it calls predictF, which learning rate and which collocation set every step uses, how long the tail
    loop runs, every line it logs.  The B200 driver must produce the same events from the same loss sequence."""
    ev = {"resample_at": [], "predictF_at": [], "lr_runs": [], "data_id_runs": []}
    st = {"t": 0, "data_id": 0, "cur": None}

    class Eng:
        n_info = 6

        def adam_init(self):
            pass

        def adam_steps(self, n, lr, want_rows=True):
            rows = np.empty((n, 6))
            for i in range(n):
                t = st["t"]
                if not ev["lr_runs"] or ev["lr_runs"][-1][1] != lr:
                    ev["lr_runs"].append([t, lr])
                if not ev["data_id_runs"] or ev["data_id_runs"][-1][1] != st["cur"]:
                    ev["data_id_runs"].append([t, st["cur"]])
                rows[i] = _synthetic_row(t)
                st["t"] = t + 1
            return rows

    class Mdl:
        engine = Eng()

        def set_data(self, data):
            st["cur"] = data

        def predict(self, zs, want_jets=False):
            ev["predictF_at"].append(st["t"])
            return np.zeros(len(zs), np.float32), np.ones(len(zs), np.float32), None

    def dataf(key, F, R, T):
        st["data_id"] += 1
        ev["resample_at"].append(st["t"])
        return st["data_id"]

    dataf.R, dataf.T = np.zeros((3, 3)), np.zeros((3, 3))
    loss = sw.adam_optimizer(dataf.R, dataf.T, Mdl(), dataf, np.ones((3, 3)), case["epoch"], sw.Key(0), lr=1e-3)
    err = capsys.readouterr().err.splitlines()
    assert st["t"] == case["n_steps"] and len(loss) == case["n_rows"]
    assert ev["resample_at"] == case["resample_at"]
    assert ev["predictF_at"] == case["predictF_at"]
    assert ev["lr_runs"] == case["lr_runs"]
    assert ev["data_id_runs"] == case["data_id_runs"]
    assert abs(float(np.sum(np.array(loss)[:, 0])) - case["loss0_checksum"]) < 1e-9 * abs(case["loss0_checksum"])
    assert err == case["stderr"]


def test_equation_front_end_never_raises_for_ui_input():
    """The reference ignores `equation` (software.py:627): None (an untouched Dash input), non-strings and
    expressions whose constants fold out of the reals must fall back to the polar Laplacian instead of
    killing the daemon training thread (callbacks/training.py:111)."""
    from pinn_based_online_pde_calculator_b200.equation import EquationError, compile_equation

    ref = sw._compile_or_reference(sw.REFERENCE_POLAR_LAPLACE, 2, "compile")
    for bad in (None, "", "   ", 17, "test equation", "u_xx + (-8)**(1/3)*u", "u_xx + (("):
        ce = sw._compile_or_reference(bad, 2, "compile")
        assert ce.ops == ref.ops and ce.consts == ref.consts, bad
    with pytest.raises(EquationError):
        compile_equation("u_xx + (-8)**(1/3)*u", d_in=2)
    with pytest.raises(EquationError):
        compile_equation(None, d_in=2)


def test_resumable_line_search_equals_direct_transcription(tmp_path):
    """The Hager-Zhang line search is a resumable state machine (csrc/lbfgs_ctl.h) so that it can run inside a
    CUDA-graph WHILE node; on the CPU it must request exactly the steps of a direct recursive transcription."""
    import shutil
    import subprocess

    if shutil.which("g++") is None:
        pytest.skip("g++ not available")
    exe = tmp_path / "ls_check"
    subprocess.run(["g++", "-O2", "-std=c++17", "-o", str(exe), os.path.join(ROOT, "tests", "cpp", "ls_coroutine_check.cpp")], check=True)
    r = subprocess.run([str(exe)], capture_output=True, text=True)
    assert r.returncode == 0 and "DIFFERENT" not in r.stdout, r.stdout
    assert r.stdout.count("same") == 8
