"""Host-side mirror of the reference's training routine, driving the CUDA engine.

Same names, argument meaning and side effects as pinn_app/software.py of the
reference ("sw:" cites): ``run_pinn_training`` (sw:626-1139) is what the Dash
callback calls (callbacks/training.py:93-105); ``adam_optimizer`` (sw:396-460),
``lbfgs_optimizer`` (sw:499-514), ``data_func_create`` (sw:521-577),
``colloc2D_set`` (sw:87-136), ``predictF`` (sw:608-623), ``gaussian2D_smooth``
(sw:71-83) keep the reference's schedule, log lines and .npz schema; the residual
loss, its gradient, Adam and L-BFGS run on the GPU through libpinn_engine.so.

Deviations (documented in DESIGN.md):
  * the ``equation`` string is compiled and used when it parses (the reference
    ignores it, sw:627); strings that do not parse fall back to the reference's
    hard-coded polar Laplacian (sw:296) with a log line;
  * jax.random / pyDOE streams are replaced by numpy streams with the same
    seed (1234, sw:685-687); the dead matplotlib figure code is not reproduced.
"""
from __future__ import annotations

import math
import os
import sys
from pathlib import Path
from typing import Callable, Dict, List, Optional, Sequence

import numpy as np

from .engine import NetworkSpec, PinnEngine
from .equation import REFERENCE_POLAR_LAPLACE, CompiledEquation, EquationError, compile_equation
from .workloads import init_params


# --------------------------------------------------------------------------- keys
class Key:
    """Stand-in for a jax PRNG key: splittable, deterministic (numpy SeedSequence)."""

    def __init__(self, entropy):
        self.ss = entropy if isinstance(entropy, np.random.SeedSequence) else np.random.SeedSequence(entropy)

    def split(self, num: int = 2) -> List["Key"]:
        return [Key(s) for s in self.ss.spawn(num)]

    def rng(self) -> np.random.Generator:
        return np.random.default_rng(self.ss)


# --------------------------------------------------------------------------- sampling (sw:71-136, 521-577)
def lhs(n: int, samples: int, rs=None) -> np.ndarray:
    """pyDOE.lhs(n, samples) (criterion=None): stratified (i+U)/N per dimension, then one independent
    permutation per dimension.  The reference draws from the GLOBAL numpy stream it seeds at sw:687 (two
    concurrent sessions clobber each other); here every run_pinn_training call owns a RandomState(seed) --
    the same legacy stream, so a single run draws exactly what the seeded global stream would give."""
    rs = rs if rs is not None else np.random
    cut = np.linspace(0, 1, samples + 1)
    u = rs.rand(samples, n)
    a, b = cut[:samples], cut[1:samples + 1]
    rd = u * (b - a)[:, None] + a[:, None]
    H = np.zeros_like(rd)
    for j in range(n):
        order = rs.permutation(range(samples))
        H[:, j] = rd[order, j]
    return H


def gaussian2D_smooth(f: np.ndarray, sig, wid) -> np.ndarray:
    """sw:71-83: normalised separable Gaussian window, 'same' 2-D convolution."""
    from scipy.signal import convolve2d

    pdf = lambda x: np.exp(-0.5 * x * x) / math.sqrt(2 * math.pi)
    xg = np.linspace(-sig[0], sig[0], int(wid[0]))
    yg = np.linspace(-sig[1], sig[1], int(wid[1]))
    window = pdf(xg) * pdf(yg)[:, None]
    return convolve2d(f, window / window.sum(), mode="same")


def colloc2D_set(key: Key, X: np.ndarray, Y: np.ndarray, F: np.ndarray, Ns: int) -> np.ndarray:
    """sw:87-136: inverse-CDF sampling of Ns points from the cell distribution F."""
    Ns = int(Ns)
    Xc, Yc, Fc = X[0:-1, 0:-1], Y[0:-1, 0:-1], F[0:-1, 0:-1]
    f = Fc.reshape(-1)
    dx = X[0, 1] - X[0, 0]
    dy = Y[1, 0] - Y[0, 0]
    seq = np.arange(f.shape[0] + 1)
    k0, k1 = key.split(2)
    b = np.hstack([0.0, np.cumsum(f)])
    c = k0.rng().uniform(size=Ns) * b[-1]
    posi_rd = np.floor(np.interp(c, b, seq))
    idx_out = np.floor(posi_rd / Fc.shape[1]).astype(np.int64)
    idx_in = (posi_rd % Fc.shape[1]).astype(np.int64)
    idx_out = np.clip(idx_out, 0, Fc.shape[0] - 1)
    posi_add = k1.rng().uniform(size=(2, Ns))
    Px = Xc[idx_out, idx_in] + posi_add[0] * dx
    Py = Yc[idx_out, idx_in] + posi_add[1] * dy
    return np.hstack((Px[:, None], Py[:, None]))


def data_func_create(N_col: Sequence[int], N_bd: int, boundary: Dict[str, float], domain: Dict[str, float],
                     sampler: str = "host", device: int = 0, rs=None):
    """sw:521-577. N_col = [n_col (LHS interior), n_bd (border-ring collocation), n_add
    (residual-adaptive)]; N_bd = points per boundary condition (sw:694).
    sampler='device' draws the collocation set with the CUDA samplers (pinn_sample_lhs /
    pinn_sample_cdf2d) and returns x_col as a CUDA tensor; boundary groups stay on the host."""
    r = np.linspace(domain["x_min"], domain["x_max"], 111)
    t = np.linspace(domain["y_min"], domain["y_max"], 111)
    R, T = np.meshgrid(r, t)
    wx = (domain["x_max"] - domain["x_min"]) / 20
    wy = (domain["y_max"] - domain["y_min"]) / 20
    inner = (R > domain["x_min"] + wx) & (R < domain["x_max"] - wx) & (T > domain["y_min"] + wy) & (T < domain["y_max"] - wy)
    F_bd = np.where(inner, 0.0, 1.0)
    span = np.array([domain["x_max"] - domain["x_min"], domain["y_max"] - domain["y_min"]])
    org = np.array([domain["x_min"], domain["y_min"]])
    num = len(boundary) // 5

    def dataf(key: Key, F, R_add, T_add):
        keys = key.split(2)
        x_bd, u_bd = [], []
        for i in range(num):
            lo = np.array([boundary[f"bd_x{i + 1}_min"], boundary[f"bd_y{i + 1}_min"]], dtype=np.float64)
            hi = np.array([boundary[f"bd_x{i + 1}_max"], boundary[f"bd_y{i + 1}_max"]], dtype=np.float64)
            x_bd.append(lhs(2, N_bd, rs) * (hi - lo) + lo)
            u_bd.append(boundary[f"bd_u{i + 1}"] * np.ones((N_bd, 1)))
        if sampler == "device":
            import torch

            from .engine import sample_cdf2d_device, sample_lhs_device

            s0, s1, s2 = (int(k.rng().integers(0, 2 ** 31)) for k in key.split(3))
            parts = [sample_lhs_device(int(N_col[0]), org, org + span, s0, device),
                     sample_cdf2d_device(int(N_col[1]), R, T, F_bd, s1, device)]
            parts += [torch.as_tensor(a, dtype=torch.float32, device=f"cuda:{device}") for a in x_bd]
            parts.append(sample_cdf2d_device(int(N_col[2]), R_add, T_add, np.asarray(F), s2, device))
            return dict(x_col=torch.cat(parts, 0), cond_bd=[x_bd, u_bd])
        x_col = lhs(2, int(N_col[0]), rs) * span + org
        xc_bd = colloc2D_set(keys[0], R, T, F_bd, N_col[1])
        xc_add = colloc2D_set(keys[1], R_add, T_add, np.asarray(F), N_col[2])
        x_col = np.vstack([x_col, xc_bd] + x_bd + [xc_add])  # BC points join the collocation set (sw:569)
        return dict(x_col=x_col, cond_bd=[x_bd, u_bd])

    dataf.R, dataf.T = R, T
    return dataf


# --------------------------------------------------------------------------- model = engine + optional frozen base
class Model:
    """The (pred_u, NN_loss) pair of the reference: sol_pred_create (sw:207) or
    mNN_pred_create (sw:221) + loss_create (sw:310), bound to one engine handle."""

    def __init__(self, net: NetworkSpec, eq: CompiledEquation, n_bc: int, lw_eqn: float, device: int = 0,
                 base: Optional["Model"] = None):
        self.net, self.eq, self.base = net, eq, base
        self.engine = PinnEngine(net, eq, n_bc=n_bc, device=device)
        self.lw, self.ref = float(lw_eqn), 1.0
        self.engine.set_loss(self.lw, self.ref)

    # f_u / f_comb (sw:213, 228) and gov_eqn (sw:283) on arbitrary points
    def _base_jets(self, z: np.ndarray):
        if self.base is None:
            return None
        return self.base.predict(z, want_jets=True)[2]

    def predict(self, z: np.ndarray, want_jets: bool = False):
        z = np.ascontiguousarray(z, dtype=np.float32)
        return self.engine.eval(z, base=self._base_jets(z), want_jets=want_jets)

    def set_data(self, data):
        x_bd = [np.ascontiguousarray(a, dtype=np.float32) for a in data["cond_bd"][0]]
        u_bd = [np.ascontiguousarray(a, dtype=np.float32).reshape(-1) for a in data["cond_bd"][1]]
        if not isinstance(data["x_col"], np.ndarray) and getattr(data["x_col"], "is_cuda", False):
            import torch  # device-resident collocation set (sampler='device')

            x_col = data["x_col"].contiguous()
            dev = lambda a: torch.as_tensor(a, dtype=torch.float32, device=x_col.device)
            base_col = self.base.engine.eval(x_col, want_u=False, want_f=False, want_jets=True)[2] if self.base is not None else None
            base_bd = [dev(self.base.predict(a)[0]) for a in x_bd] if self.base is not None else None
            self.engine.set_points(x_col, [dev(a) for a in x_bd], [dev(a) for a in u_bd], base_col=base_col, base_bd=base_bd)
            return
        x_col = np.ascontiguousarray(data["x_col"], dtype=np.float32)
        base_col = self._base_jets(x_col)
        base_bd = None
        if self.base is not None:
            base_bd = [self.base.predict(a)[0] for a in x_bd]
        self.engine.set_points(x_col, x_bd, u_bd, base_col=base_col, base_bd=base_bd)

    def set_ref(self, ref: float):
        """NN_loss.ref = ... (sw:739)."""
        self.ref = float(ref)
        self.engine.set_loss(self.lw, self.ref)

    def loss_info(self) -> np.ndarray:
        return self.engine.loss_grad(want_grad=False)[1]

    def close(self):
        self.engine.close()


def predictF(model: Model, z1: np.ndarray, z2: np.ndarray) -> np.ndarray:
    """sw:608-623: residual map f^2/mean(f^2)+0.5, 5x5 Gaussian smoothing."""
    z_star = np.hstack((z1.reshape(-1)[:, None], z2.reshape(-1)[:, None]))
    _, f0, _ = model.predict(z_star)
    f_sq = f0.astype(np.float64) ** 2
    f_nm = f_sq / np.mean(f_sq) + 0.5
    return gaussian2D_smooth(f_nm.reshape(z1.shape), [1, 1], [5, 5])


# --------------------------------------------------------------------------- optimisers
def _log_line(step, info) -> str:
    return (f"Step: {step} | Loss: {info[0]:.4e} | Loss_d: {info[1]:.4e} | Loss_e: {info[2]:.4e} | ")


def adam_optimizer(R_add, T_add, model: Model, dataf, F, epoch: int, key_adam: Key, lr: float = 1e-3):
    """sw:396-460 with the per-step work on the GPU.  Steps between host events
    (resample every 100, predictF every 2000, LR test every 4000) are replayed as a
    CUDA graph without host synchronisation."""
    eng = model.engine
    eng.adam_init()
    loss_all: List[np.ndarray] = []
    key = key_adam
    model.set_data(dataf(key, F, R_add, T_add))
    R, T = dataf.R, dataf.T
    epoch = int(epoch)
    nc = int(np.round(epoch / 5))
    nc0 = 2000
    step, info = 0, None
    while step < epoch:
        # next step index (inclusive) after which the host must act
        c1 = ((step + 99) // 100) * 100 or 100          # s % 100 == 0 and s > 0
        c2 = ((step + nc0) // nc0) * nc0 - 1            # (s + 1) % nc0 == 0
        nxt = min(c1, c2, epoch - 1)
        s = nxt
        resample = s % 100 == 0 and s > 0
        # The steps are enqueued without waiting; the host samples the NEXT collocation set (sw:420-422: it depends on
        # the key and on F, which only changes at the predictF boundaries that end their own chunk) while they run.
        begun = getattr(eng, "adam_steps_begin", None) is not None and eng.adam_steps_begin(nxt - step + 1, lr)
        next_data = None
        if begun and resample:
            key = key.split(1)[0]
            next_data = dataf(key, F, R_add, T_add)
        rows = eng.adam_steps_end(nxt - step + 1) if begun else eng.adam_steps(nxt - step + 1, lr)
        loss_all.extend(rows)
        info = rows[-1]
        if resample:
            print(_log_line(s, info), file=sys.stderr)
            if next_data is None:
                key = key.split(1)[0]
                next_data = dataf(key, F, R_add, T_add)
            model.set_data(next_data)
        if (s + 1) % nc0 == 0:
            F = predictF(model, R, T)
        if (s + 1) % (2 * nc0) == 0:
            lossend = np.array(loss_all[-2 * nc0:])[:, 0]
            lc1, lc2 = lossend[0:nc0], lossend[nc0:]
            mm12 = abs(np.mean(lc1) - np.mean(lc2))
            stdl2 = np.std(lc2)
            if mm12 / stdl2 < 0.4:
                lr = lr / 2  # new optax.adam(lr) with the OLD opt_state (sw:439-440)
            print(f"learning rate for Adam: {lr:.4e} | mean: {mm12:.3e} | std: {stdl2:.3e}", file=sys.stderr)
        step = s + 1
    lossend = np.array(loss_all[-nc:])[:, 0] if nc > 0 else np.array(loss_all)[:, 0]
    lmin, llast = np.min(lossend), lossend[-1]
    for _ in range(2 * nc0):  # sw:450-456
        if llast < lmin:
            break
        rows = eng.adam_steps(1, lr)
        info = rows[-1]
        llast = info[0]
        loss_all.append(info)
    print(_log_line(epoch - 1, info), file=sys.stderr)
    return loss_all


def lbfgs_optimizer(model: Model, epoch: int, value_unnormalised: bool = True):
    """sw:499-514: max_iterations = int32(epoch/3), tolerance 1e-10; one loss_info row
    and one 'Step: NaN' line per objective evaluation (sw:485-488)."""
    max_iter = int(np.int32(epoch / 3))

    def on_eval(x):
        print(f"Step: NaN | Loss: {x[0]:.4e} | Loss_d: {x[1]:.4e} | Loss_e: {x[2]:.4e}")

    res, rows = model.engine.lbfgs(max_iter, 1e-10, value_unnormalised, on_eval)
    print(f" Total iterations: {res['evaluations']}")
    return rows, res


# --------------------------------------------------------------------------- driver (sw:626-1139)
def _compile_or_reference(equation, d_in: int, mode: str) -> CompiledEquation:
    """The reference ignores `equation` entirely (sw:627), so NOTHING the UI can put there may kill the
    training thread: None (an untouched Dash input), non-strings and expressions whose constant folding
    leaves the reals all fall back to the reference's hard-coded polar Laplacian."""
    if mode != "reference" and isinstance(equation, str) and equation.strip():
        try:
            return compile_equation(equation, d_in=d_in)
        except Exception as e:  # EquationError, or anything a malformed expression provokes in the front end
            print(f"equation {equation!r} not compiled ({type(e).__name__}: {e}); using the reference's polar Laplacian",
                  file=sys.stderr)
    elif mode != "reference":
        print(f"equation {equation!r} is empty; using the reference's polar Laplacian", file=sys.stderr)
    return compile_equation(REFERENCE_POLAR_LAPLACE, d_in=d_in)


def _exact_default(R, T):
    return np.log(R) / np.log(0.1)  # sw:815


def run_pinn_training(
        equation: str,
        boundary: dict,
        domain: dict,
        scl: float,
        epsil: float,
        sample_points: dict,
        network_size: dict,
        testing_size: dict,
        epochs: dict,
        equation_weight: dict,
        output_dir: str,
        *,
        feature_map: str = "polar",
        exact_solution: Optional[Callable] = None,
        equation_mode: Optional[str] = None,
        stage2: bool = True,
        device: int = 0,
        seed: int = 1234,
        n_bd_points: int = 100,
        sampler: str = "auto",
):
    """Same 11 kwargs as the reference (sw:626-638); keyword-only extras are extensions."""
    m_x_min, m_x_max = domain["x_min"], domain["x_max"]
    m_y_min, m_y_max = domain["y_min"], domain["y_max"]
    m_depth, m_width = network_size["depth"], network_size["width"]  # depth = units, width = layers (sw:712)
    m_nx, m_ny = int(testing_size["x"]), int(testing_size["y"])
    m_adam, m_lbfgs = epochs["adam"], epochs["lbfgs"]
    m_f, m_df = equation_weight["f"], equation_weight["df"]
    mode = equation_mode or os.environ.get("PINN_B200_EQUATION_MODE", "compile")
    exact = exact_solution or _exact_default

    if sampler == "auto":  # host numpy streams for reference-sized sets, CUDA samplers for large ones
        sampler = "device" if sample_points["n_col"] >= 200_000 else "host"
    as_np = lambda a: a if isinstance(a, np.ndarray) else a.cpu().numpy()
    base_dir = Path(output_dir)
    base_dir.mkdir(parents=True, exist_ok=True)

    key = Key(seed)
    rs = np.random.RandomState(seed)  # per-call stream (the reference seeds the process-global one, sw:687)
    keys = key.split(10)
    N_col = np.array([sample_points["n_col"], sample_points["n_bd"], sample_points["n_add"]])
    N_bd = n_bd_points

    r = np.linspace(m_x_min, m_x_max, m_nx)
    t = np.linspace(m_y_min, m_y_max, m_ny)
    R, T = np.meshgrid(r, t)
    X_star = np.hstack((R.reshape(-1)[:, None], T.reshape(-1)[:, None]))
    lb, ub = [m_x_min, m_y_min], [m_x_max, m_y_max]
    n_bc = len(boundary) // 5
    eq = _compile_or_reference(equation, 2, mode)
    limit4 = [domain["x_min"], domain["x_max"], domain["y_min"], domain["y_max"]]

    def save_colpoints(U, X_col, path):  # colpoint_plot's np.savez (sw:600-605)
        np.savez(path, U=U, X_col=X_col, limit=np.array(limit4))

    # ---------------- stage 1
    net1 = NetworkSpec(n_hidden=int(m_width), width=int(m_depth), lb=lb, ub=ub, scl=scl, epsil=epsil, act_first=0,
                       feature_map=feature_map, d_in=2)
    model1 = Model(net1, eq, n_bc, lw_eqn=m_f, device=device)
    model1.engine.set_params(init_params(net1, seed))
    dataf1 = data_func_create(N_col, N_bd, boundary, domain, sampler=sampler, device=device, rs=rs)
    Fs = R * 0 + 1
    key_adam = keys[1]
    key_lbfgs = keys[2].split(1)
    Rg, Tg = dataf1.R, dataf1.T
    Fg = Rg * 0 + 1
    data1 = dataf1(key_adam, Fg, Rg, Tg)
    save_colpoints(Fs, as_np(data1["x_col"]), base_dir / "collocation_point_1.npz")

    model1.set_data(data1)
    model1.set_ref(1.0)
    model1.set_ref(model1.loss_info()[0])  # sw:738-739

    loss1 = adam_optimizer(Rg, Tg, model1, dataf1, Fg, m_adam, key_adam, lr=1e-3)
    Fg = predictF(model1, Rg, Tg)
    model1.set_data(dataf1(key_lbfgs[0], Fg, Rg, Tg))
    loss2, _ = lbfgs_optimizer(model1, m_lbfgs)
    Fg = predictF(model1, Rg, Tg)

    u_p1, f_p1, _ = model1.predict(X_star)
    U = u_p1.astype(np.float64).reshape(R.shape)
    F = f_p1.astype(np.float64).reshape(R.shape)
    loss_all1 = np.array(list(loss1) + list(loss2))
    np.savez(base_dir / "solution_residual_1.npz", r=r, t_vec=t, U=U, F=F)  # sw:806-811
    U_real = exact(R, T)
    Error = U - U_real
    np.savez(base_dir / "error_1.npz", r=R[0, :], t=T[:, 0], Error=Error)  # sw:829-834
    np.savez(base_dir / "loss_1.npz", loss=loss_all1)  # sw:866
    np.savez(base_dir / "boundary_loss_1.npz", loss_xy_l=loss_all1[:, 3],
             loss_xy_r=loss_all1[:, 4] if loss_all1.shape[1] > 4 else loss_all1[:, 3])  # sw:890-897
    r1_rms = float(np.sqrt(np.mean(F ** 2)))  # sw:899-903
    e1_rms = float(np.sqrt(np.mean(Error ** 2)))
    diff = r1_rms / e1_rms
    mag = np.abs(np.fft.fftshift(np.fft.fft2(F)))  # sw:906-936
    np.savez(base_dir / "frequency_spectrum.npz",
             freq_x=np.fft.fftshift(np.fft.fftfreq(m_nx, d=(R[0, 1] - R[0, 0]))),
             freq_t=np.fft.fftshift(np.fft.fftfreq(m_ny, d=(T[1, 0] - T[0, 0]))), log_mag=np.log1p(mag))
    result = dict(loss_1=loss_all1, U1=U, F1=F, Error1=Error, r1_rms=r1_rms, e1_rms=e1_rms)
    if not stage2:
        model1.close()
        return result

    # ---------------- stage 2 (sw:938-1139): 6x50, sin first layer, trained on the stage-1 remainder
    n_hl2, n_unit2 = 6, 50
    scl2 = 30.0 if e1_rms > 50 else diff
    lw2 = m_f / diff
    epsil2 = e1_rms
    net2 = NetworkSpec(n_hidden=n_hl2, width=n_unit2, lb=lb, ub=ub, scl=scl2, epsil=epsil2, act_first=1,
                       feature_map=feature_map, d_in=2)
    model2 = Model(net2, eq, n_bc, lw_eqn=lw2, device=device, base=model1)
    model2.engine.set_params(init_params(net2, seed + 3))
    dataf2 = data_func_create(N_col * 2, N_bd * 2, boundary, domain, sampler=sampler, device=device, rs=rs)
    key_adam = keys[4]
    key_lbfgs = keys[5].split(1)
    Fg = Rg * 0 + 1
    data2 = dataf2(key_adam, Fg, Rg, Tg)
    save_colpoints(Fs, as_np(data2["x_col"]), base_dir / "collocation_point_2.npz")
    model2.set_data(data2)
    model2.set_ref(1.0)
    model2.set_ref(model2.loss_info()[0])
    loss1b = adam_optimizer(Rg, Tg, model2, dataf2, Fg, m_adam * 3, key_adam, lr=1e-3)
    Fg = predictF(model2, Rg, Tg)
    model2.set_data(dataf2(key_lbfgs[0], Fg, Rg, Tg))
    loss2b, _ = lbfgs_optimizer(model2, m_lbfgs * 3)

    u_p2, f_p2, _ = model2.predict(X_star)
    U2 = u_p2.astype(np.float64).reshape(R.shape)
    F2 = f_p2.astype(np.float64).reshape(R.shape)
    loss_all2 = np.array(list(loss1b) + list(loss2b))
    loss_all = np.vstack([loss_all1, loss_all2])
    np.savez(base_dir / "solution_residual_2.npz", r=R[:, 0], t=T[0, :], U=U2, F=F2)  # sw:1037-1046
    Error2 = U2 - U_real
    np.savez(base_dir / "error_2.npz", r=R[0, :], t=T[:, 0], Error=Error2)
    np.savez(base_dir / "loss_2.npz", loss=loss_all)
    np.savez(base_dir / "boundary_loss_2.npz", loss_xy_l=loss_all[:, 3],
             loss_xy_r=loss_all[:, 4] if loss_all.shape[1] > 4 else loss_all[:, 3])
    result.update(loss_2=loss_all, U2=U2, F2=F2, Error2=Error2)
    model2.close()
    model1.close()
    return result
