"""ctypes binding of libpinn_engine.so (include/pinn_engine.h) and the host-side
mirror of the reference's operator API for the hot path:

    loss_fun(params, data) -> (loss_n, loss_info)         software.py:318
    grad(lossf, has_aux=True)(params, data)                software.py:390, 479
    adam_minimizer(lossf, params, data, opt, opt_state)    software.py:388
    f(params_1d) -> (loss_value, grads_1d)                 software.py:475
    f_u(params, z) -> u ; gov_eqn(f_u, z) -> f             software.py:213, 283

There is NO CPU fallback: importing works anywhere (so CPU-only tests can check
the exported symbols), but creating an engine without the library or without a
CUDA device raises.
"""
from __future__ import annotations

import ctypes as C
import os
from dataclasses import dataclass, field
from typing import List, Optional, Sequence

import numpy as np

from .equation import CompiledEquation

_LIB_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "libpinn_engine.so")
_lib = None

EXPORTS = [
    "pinn_last_error", "pinn_device_count", "pinn_engine_create", "pinn_engine_destroy",
    "pinn_engine_set_stream", "pinn_engine_num_params", "pinn_engine_num_loss_info",
    "pinn_engine_tile_points", "pinn_engine_launches_per_eval", "pinn_engine_launches_per_adam_step", "pinn_engine_set_params",
    "pinn_engine_get_params", "pinn_engine_set_points", "pinn_engine_set_global_counts",
    "pinn_engine_set_loss", "pinn_engine_loss_grad", "pinn_engine_adam_init", "pinn_engine_adam_steps", "pinn_engine_adam_rows",
    "pinn_engine_eval", "pinn_engine_lbfgs", "pinn_nccl_unique_id", "pinn_engine_init_nccl",
    "pinn_fma_peak", "pinn_engine_last_ms", "pinn_engine_time_kernels", "pinn_engine_kernel_kind", "pinn_engine_phase_profile", "pinn_sample_lhs", "pinn_sample_cdf2d", "pinn_engine_sync", "pinn_umma_probe", "pinn_engine_prefetch_points", "pinn_engine_commit_points", "pinn_engine_umma_clocks", "pinn_lbfgs_direction_test",
    "pinn_engine_lbfgs_trace", "pinn_engine_lbfgs_trace_rows", "pinn_engine_lbfgs_trace_get", "pinn_engine_lbfgs_host_syncs",
]


class PinnSpecC(C.Structure):
    _fields_ = [
        ("d_in", C.c_int32), ("feat_mode", C.c_int32), ("n_hidden", C.c_int32), ("width", C.c_int32),
        ("act_first", C.c_int32), ("act_hidden", C.c_int32), ("scl", C.c_float), ("epsil", C.c_float),
        ("lb", C.c_float * 3), ("ub", C.c_float * 3), ("n1", C.c_int32), ("n2", C.c_int32), ("mix", C.c_int32),
        ("n_ops", C.c_int32), ("ops", C.POINTER(C.c_int32)), ("n_consts", C.c_int32),
        ("consts", C.POINTER(C.c_float)), ("n_aux_col", C.c_int32), ("n_bc", C.c_int32),
        ("n_aux_user", C.c_int32), ("n_aux_ops", C.c_int32), ("aux_ops", C.POINTER(C.c_int32)),
        ("lap_beta", C.c_float * 3), ("lap_aux", C.c_int32 * 3),
    ]


class LbfgsResultC(C.Structure):
    _fields_ = [("iterations", C.c_int32), ("evaluations", C.c_int32), ("converged", C.c_int32),
                ("failed", C.c_int32), ("final_loss", C.c_double)]


EVAL_CB = C.CFUNCTYPE(None, C.POINTER(C.c_double), C.c_int32, C.c_void_p)


def load_library(path: Optional[str] = None):
    """dlopen the C-ABI library and declare signatures. Raises if it is missing."""
    global _lib
    if _lib is not None and path is None:
        return _lib
    p = path or os.environ.get("PINN_B200_LIB") or _LIB_PATH   # PINN_B200_LIB: alternative build (kernel experiments)
    if not os.path.exists(p):
        raise RuntimeError(
            f"{p} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(there is no CPU fallback for the B200 engine)")
    lib = C.CDLL(p)
    lib.pinn_last_error.restype = C.c_char_p
    lib.pinn_device_count.restype = C.c_int
    lib.pinn_engine_create.argtypes = [C.POINTER(PinnSpecC), C.c_int, C.POINTER(C.c_void_p)]
    lib.pinn_engine_destroy.argtypes = [C.c_void_p]
    lib.pinn_engine_destroy.restype = None
    lib.pinn_engine_set_stream.argtypes = [C.c_void_p, C.c_void_p]
    lib.pinn_engine_sync.argtypes = [C.c_void_p]
    lib.pinn_engine_prefetch_points.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p]
    lib.pinn_engine_commit_points.argtypes = [C.c_void_p]
    lib.pinn_engine_num_params.argtypes = [C.c_void_p]
    lib.pinn_engine_num_params.restype = C.c_int64
    lib.pinn_engine_num_loss_info.argtypes = [C.c_void_p]
    lib.pinn_engine_tile_points.argtypes = [C.c_void_p]
    lib.pinn_engine_launches_per_eval.argtypes = [C.c_void_p]
    lib.pinn_engine_launches_per_adam_step.argtypes = [C.c_void_p]
    lib.pinn_engine_kernel_kind.argtypes = [C.c_void_p]
    lib.pinn_engine_phase_profile.argtypes = [C.c_void_p, C.POINTER(C.c_int64)]
    lib.pinn_sample_lhs.argtypes = [C.c_int, C.c_void_p, C.c_uint32, C.c_int64, C.c_int32, C.POINTER(C.c_float),
                                    C.POINTER(C.c_float), C.c_void_p, C.c_int32, C.c_int32]
    lib.pinn_sample_cdf2d.argtypes = [C.c_int, C.c_void_p, C.c_uint32, C.c_int64, C.c_void_p, C.c_int32, C.c_int32,
                                      C.c_float, C.c_float, C.c_float, C.c_float, C.c_void_p, C.c_int32]
    lib.pinn_engine_set_params.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
    lib.pinn_engine_get_params.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
    lib.pinn_engine_set_points.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_int32,
                                           C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), C.POINTER(C.c_void_p),
                                           C.POINTER(C.c_int64), C.c_int]
    lib.pinn_engine_set_global_counts.argtypes = [C.c_void_p, C.c_int64, C.POINTER(C.c_int64)]
    lib.pinn_engine_set_loss.argtypes = [C.c_void_p, C.c_double, C.c_double]
    lib.pinn_engine_loss_grad.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    lib.pinn_engine_adam_init.argtypes = [C.c_void_p]
    lib.pinn_engine_adam_steps.argtypes = [C.c_void_p, C.c_int32, C.c_double, C.c_void_p]
    lib.pinn_engine_adam_rows.argtypes = [C.c_void_p, C.c_int32, C.c_void_p]
    lib.pinn_engine_eval.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p,
                                     C.c_void_p, C.c_void_p, C.c_int]
    lib.pinn_engine_lbfgs.argtypes = [C.c_void_p, C.c_int32, C.c_double, C.c_int32, EVAL_CB, C.c_void_p,
                                      C.POINTER(LbfgsResultC)]
    lib.pinn_engine_lbfgs_trace.argtypes = [C.c_void_p, C.c_int32]
    lib.pinn_engine_lbfgs_trace_rows.argtypes = [C.c_void_p]
    lib.pinn_engine_lbfgs_trace_get.argtypes = [C.c_void_p, C.c_void_p, C.c_int32]
    lib.pinn_engine_lbfgs_host_syncs.argtypes = [C.c_void_p]
    lib.pinn_nccl_unique_id.argtypes = [C.c_void_p]
    lib.pinn_engine_init_nccl.argtypes = [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32]
    lib.pinn_fma_peak.argtypes = [C.c_int, C.c_int, C.POINTER(C.c_double)]
    lib.pinn_engine_last_ms.argtypes = [C.c_void_p]
    lib.pinn_engine_last_ms.restype = C.c_double
    lib.pinn_engine_time_kernels.argtypes = [C.c_void_p, C.c_int32, C.c_int64, C.POINTER(C.c_double),
                                             C.POINTER(C.c_double)]
    if path is None:
        _lib = lib
    return lib


def _check(lib, rc: int):
    if rc != 0:
        raise RuntimeError("pinn_engine: " + (lib.pinn_last_error() or b"unknown error").decode())


@dataclass
class NetworkSpec:
    """sol_pred_create / neural_net arguments (software.py:158-218)."""
    n_hidden: int
    width: int
    lb: Sequence[float]
    ub: Sequence[float]
    scl: float = 1.0
    epsil: float = 1.0
    act_first: int = 0          # 0 tanh, 1 sin (software.py:170)
    act_hidden: int = 0         # extension: all-layer sin
    feature_map: str = "polar"  # 'polar' (reference, software.py:172-175) | 'affine'
    d_in: int = 2

    @property
    def n_feat(self) -> int:
        return 3 if self.feature_map == "polar" else self.d_in

    @property
    def layer_widths(self) -> List[int]:
        return [self.n_feat] + self.n_hidden * [self.width] + [1]

    @property
    def n_params(self) -> int:
        lw = self.layer_widths
        return sum(a * b + b for a, b in zip(lw[:-1], lw[1:]))


def _ptr(a):
    """Raw pointer of a numpy array (host) or a torch CUDA tensor (device)."""
    if a is None:
        return None
    if isinstance(a, np.ndarray):
        return a.ctypes.data_as(C.c_void_p)
    return C.c_void_p(a.data_ptr())


def _is_device(a) -> bool:
    return a is not None and not isinstance(a, np.ndarray) and getattr(a, "is_cuda", False)


def _as_f32(a):
    if a is None:
        return None
    if isinstance(a, np.ndarray):
        return np.ascontiguousarray(a, dtype=np.float32)
    import torch

    if torch.is_tensor(a):
        if a.is_cuda:
            return a.contiguous().float() if a.dtype != torch.float32 or not a.is_contiguous() else a
        return np.ascontiguousarray(a.detach().cpu().numpy(), dtype=np.float32)
    return np.ascontiguousarray(np.asarray(a), dtype=np.float32)


class PinnEngine:
    """One engine handle = one (network, residual, loss) closure of the reference."""

    def __init__(self, net: NetworkSpec, eq: CompiledEquation, n_bc: int, device: int = 0, lib_path: Optional[str] = None):
        self.lib = load_library(lib_path)
        if self.lib.pinn_device_count() <= 0:
            raise RuntimeError("pinn_engine: no CUDA device (the B200 engine has no CPU fallback)")
        if eq.d_in != net.d_in:
            raise ValueError("equation and network disagree on d_in")
        self.net, self.eq, self.n_bc, self.device = net, eq, n_bc, device
        spec = PinnSpecC()
        spec.d_in = net.d_in
        spec.feat_mode = 1 if net.feature_map == "polar" else 0
        spec.n_hidden, spec.width = net.n_hidden, net.width
        spec.act_first, spec.act_hidden = net.act_first, net.act_hidden
        spec.scl, spec.epsil = float(net.scl), float(net.epsil)
        for i in range(3):
            spec.lb[i] = float(net.lb[i]) if i < len(net.lb) else 0.0
            spec.ub[i] = float(net.ub[i]) if i < len(net.ub) else 1.0
        spec.n1, spec.n2, spec.mix = eq.n1, eq.n2, eq.mix
        self._ops = (C.c_int32 * len(eq.ops))(*eq.ops)
        self._consts = (C.c_float * max(1, len(eq.consts)))(*(eq.consts or [0.0]))
        spec.n_ops, spec.ops = len(eq.ops), self._ops
        spec.n_consts, spec.consts = len(eq.consts), self._consts
        spec.n_aux_col, spec.n_bc = eq.n_aux, n_bc
        self._aux_ops = (C.c_int32 * max(1, len(eq.aux_ops)))(*(eq.aux_ops or [0]))
        spec.n_aux_user, spec.n_aux_ops, spec.aux_ops = eq.n_aux_user, len(eq.aux_ops), self._aux_ops
        for i in range(3):
            spec.lap_beta[i] = float(eq.lap_beta[i])
            spec.lap_aux[i] = int(eq.lap_aux[i])
        h = C.c_void_p()
        _check(self.lib, self.lib.pinn_engine_create(C.byref(spec), device, C.byref(h)))
        self.h = h
        self.n_params = int(self.lib.pinn_engine_num_params(h))
        self.n_info = int(self.lib.pinn_engine_num_loss_info(h))
        self.K = eq.K
        self.kernel = ("simt_fp32", "mma_3xtf32", "umma_3xtf32", "tc_bf16x3")[int(self.lib.pinn_engine_kernel_kind(h))]
        self._keep = []  # device tensors borrowed by the engine
        self.lref = 1.0
        self.lw = 1.0

    # ---- lifetime
    def close(self):
        if getattr(self, "h", None):
            self.lib.pinn_engine_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_stream(self, cuda_stream_ptr: int):
        _check(self.lib, self.lib.pinn_engine_set_stream(self.h, C.c_void_p(cuda_stream_ptr)))
        self._shared_stream = True

    def prefetch_points(self, x_col, x_bd: Sequence = (), u_bd: Sequence = ()):
        """Start copying the NEXT point set (host arrays, same shapes as the current one; pinned memory
        makes the copy asynchronous) while the step in flight computes; see commit_points."""
        x_col = _as_f32(x_col)
        xb = [_as_f32(a) for a in x_bd]
        ub = [_as_f32(a).reshape(-1) for a in u_bd]
        n = len(xb)
        PX = (C.c_void_p * max(1, n))(*[_ptr(a) for a in xb])
        PU = (C.c_void_p * max(1, n))(*[_ptr(a) for a in ub])
        NB = (C.c_int64 * max(1, n))(*[int(a.shape[0]) for a in xb])
        # keep the host buffers alive while their asynchronous copy may be in flight (two generations)
        self._prefetched = (getattr(self, "_prefetched", (None,))[-1], (x_col, xb, ub))
        _check(self.lib, self.lib.pinn_engine_prefetch_points(self.h, _ptr(x_col), int(x_col.shape[0]), n, PX, PU, NB))

    def commit_points(self):
        """Swap the prefetched point set in (engine stream; does not block)."""
        _check(self.lib, self.lib.pinn_engine_commit_points(self.h))

    def sync(self):
        _check(self.lib, self.lib.pinn_engine_sync(self.h))

    def _torch_inputs_ready(self):
        """Device tensors handed to the engine were produced on torch's current stream; unless the engine
        shares that stream (set_stream), wait for it so the engine's own stream sees complete data."""
        if not getattr(self, "_shared_stream", False):
            import torch

            torch.cuda.current_stream(self.device).synchronize()

    # ---- parameters (ravel_pytree order, software.py:466)
    def set_params(self, flat):
        a = _as_f32(flat)
        if _is_device(a):
            self._torch_inputs_ready()
        n = a.size if isinstance(a, np.ndarray) else a.numel()
        if n != self.n_params:
            raise ValueError(f"expected {self.n_params} parameters, got {n}")
        _check(self.lib, self.lib.pinn_engine_set_params(self.h, _ptr(a), int(_is_device(a))))
        if _is_device(a):
            # the device-to-device copy runs on the engine's own stream: keep the (possibly temporary) source
            # alive until it has completed, otherwise torch's caching allocator may hand it out again
            self._param_src = a
            if not getattr(self, "_shared_stream", False):
                self.sync()
                self._param_src = None

    def get_params(self) -> np.ndarray:
        out = np.empty(self.n_params, dtype=np.float32)
        _check(self.lib, self.lib.pinn_engine_get_params(self.h, _ptr(out), 0))
        return out

    # ---- data dict (software.py:572)
    def set_points(self, x_col, x_bd: Sequence = (), u_bd: Sequence = (), aux_col=None, base_col=None,
                   base_bd: Optional[Sequence] = None):
        x_col = _as_f32(x_col)
        dev = _is_device(x_col)
        aux_col, base_col = _as_f32(aux_col), _as_f32(base_col)
        xb = [_as_f32(a) for a in x_bd]
        ub = [_as_f32(a).reshape(-1) for a in u_bd]
        bb = [(_as_f32(a).reshape(-1) if a is not None else None) for a in (base_bd or [None] * len(xb))]
        if dev:
            for a in [aux_col, base_col] + xb + ub + bb:
                if a is not None and not _is_device(a):
                    raise ValueError("mixing host and device buffers in set_points")
        n_col = x_col.shape[0]
        n = len(xb)
        if dev:
            self._torch_inputs_ready()
        PX = (C.c_void_p * max(1, n))(*[_ptr(a) for a in xb])
        PU = (C.c_void_p * max(1, n))(*[_ptr(a) for a in ub])
        PB = (C.c_void_p * max(1, n))(*[_ptr(a) for a in bb])
        NB = (C.c_int64 * max(1, n))(*[int(a.shape[0]) for a in xb])
        self._keep = [x_col, aux_col, base_col]
        _check(self.lib, self.lib.pinn_engine_set_points(self.h, _ptr(x_col), n_col, _ptr(aux_col), _ptr(base_col), n,
                                                         PX, PU, PB, NB, int(dev)))
        self.n_col = n_col
        self.n_bd = [int(a.shape[0]) for a in xb]

    def set_global_counts(self, n_col_global: int, n_bd_global: Sequence[int]):
        NB = (C.c_int64 * max(1, len(n_bd_global)))(*[int(v) for v in n_bd_global])
        _check(self.lib, self.lib.pinn_engine_set_global_counts(self.h, int(n_col_global), NB))

    def set_loss(self, lw_eqn: float, lref: float):
        """loss_fun.lw[0], loss_fun.ref (software.py:381-382)."""
        self.lw, self.lref = float(lw_eqn), float(lref)
        _check(self.lib, self.lib.pinn_engine_set_loss(self.h, self.lw, self.lref))

    # ---- grad(lossf, has_aux=True) (software.py:390)
    def loss_grad(self, params=None, want_grad: bool = True):
        """Returns (grads_flat float32 [P] of loss/lref, loss_info float64 [3+n_bc+1] un-normalised)."""
        import torch

        info = np.empty(self.n_info, dtype=np.float64)
        g = torch.empty(self.n_params, dtype=torch.float32, device=f"cuda:{self.device}") if want_grad else None
        p = None
        if params is not None:
            p = params if _is_device(params) else torch.as_tensor(np.asarray(params, dtype=np.float32)).to(f"cuda:{self.device}")
            self._torch_inputs_ready()
        _check(self.lib, self.lib.pinn_engine_loss_grad(self.h, _ptr(p), _ptr(g), _ptr(info)))
        return g, info

    def loss_fun(self, params=None):
        """loss_fun(params, data) -> (loss_n, loss_info) (software.py:318-379)."""
        _, info = self.loss_grad(params, want_grad=False)
        return info[0] / self.lref, info

    # ---- adam_minimizer (software.py:387-393)
    def adam_init(self):
        _check(self.lib, self.lib.pinn_engine_adam_init(self.h))

    def adam_steps(self, n_steps: int, lr: float, want_rows: bool = True) -> Optional[np.ndarray]:
        rows = np.empty((n_steps, self.n_info), dtype=np.float64) if want_rows else None
        _check(self.lib, self.lib.pinn_engine_adam_steps(self.h, int(n_steps), float(lr), _ptr(rows)))
        return rows

    RING_CAP = 4096

    def adam_steps_begin(self, n_steps: int, lr: float) -> bool:
        """Enqueue n_steps Adam steps WITHOUT waiting for them (software.py:416-425: the caller samples the next collocation
        set meanwhile) -- False when the rows would not fit the device ring (use adam_steps then)."""
        if n_steps > self.RING_CAP:
            return False
        _check(self.lib, self.lib.pinn_engine_adam_steps(self.h, int(n_steps), float(lr), None))
        return True

    def adam_steps_end(self, n_steps: int) -> np.ndarray:
        """loss_info rows of the steps adam_steps_begin enqueued (synchronises)."""
        rows = np.empty((n_steps, self.n_info), dtype=np.float64)
        _check(self.lib, self.lib.pinn_engine_adam_rows(self.h, int(n_steps), _ptr(rows)))
        return rows

    def launches_per_eval(self) -> int:
        """kernels one loss/gradient evaluation enqueues (an Adam step adds one)"""
        return int(self.lib.pinn_engine_launches_per_eval(self.h))

    def launches_per_adam_step(self) -> int:
        """kernels one Adam step of adam_steps enqueues (evaluation kernels + the fused tail kernel)"""
        return int(self.lib.pinn_engine_launches_per_adam_step(self.h))

    def last_ms(self) -> float:
        return float(self.lib.pinn_engine_last_ms(self.h))

    def time_kernels(self, reps: int = 5, flush_bytes: int = 256 << 20):
        """(col_ms, bc_ms): average device time of each fused kernel launched alone."""
        a, b = C.c_double(), C.c_double()
        _check(self.lib, self.lib.pinn_engine_time_kernels(self.h, reps, flush_bytes, C.byref(a), C.byref(b)))
        return a.value, b.value

    def phase_profile(self):
        """clock64 totals of CTA 0 per phase of the (tensor-core) collocation kernel."""
        out = (C.c_int64 * 8)()
        _check(self.lib, self.lib.pinn_engine_phase_profile(self.h, out))
        names = ["fwd_gemm", "act_fwd", "output_residual", "act_bwd", "restage", "wgrad", "dgrad", "rest"]
        return dict(zip(names, [int(v) for v in out]))

    # ---- f_u / gov_eqn (software.py:213, 283)
    def eval(self, z, aux=None, base=None, want_u=True, want_f=True, want_jets=False):
        z = _as_f32(z)
        n = z.shape[0]
        dev = _is_device(z)
        if dev:
            import torch

            mk = lambda *s: torch.empty(*s, dtype=torch.float32, device=z.device)
        else:
            mk = lambda *s: np.empty(s, dtype=np.float32)
        u = mk(n) if want_u else None
        f = mk(n) if want_f else None
        j = mk(n, self.K) if want_jets else None
        if dev:
            self._torch_inputs_ready()
        _check(self.lib, self.lib.pinn_engine_eval(self.h, _ptr(z), n, _ptr(_as_f32(aux)), _ptr(_as_f32(base)),
                                                   _ptr(u), _ptr(f), _ptr(j), int(dev)))
        if dev and not getattr(self, "_shared_stream", False):
            self.sync()  # outputs are consumed on torch's stream
        return u, f, j

    # ---- lbfgs_optimizer (software.py:499-514)
    def lbfgs(self, max_iter: int, tol: float = 1e-10, value_unnormalised: bool = True, on_eval=None):
        res = LbfgsResultC()
        rows: List[np.ndarray] = []

        def _cb(ptr, n, user):
            row = np.array([ptr[i] for i in range(n)], dtype=np.float64)
            rows.append(row)
            if on_eval is not None:
                on_eval(row)

        cb = EVAL_CB(_cb)
        _check(self.lib, self.lib.pinn_engine_lbfgs(self.h, int(max_iter), float(tol), int(value_unnormalised), cb,
                                                    None, C.byref(res)))
        return dict(iterations=res.iterations, evaluations=res.evaluations, converged=bool(res.converged),
                    failed=bool(res.failed), final_loss=res.final_loss), rows

    def lbfgs_trace(self, cap: int):
        """test hook: keep the first `cap` trial parameter vectors of the following lbfgs() calls"""
        _check(self.lib, self.lib.pinn_engine_lbfgs_trace(self.h, int(cap)))

    def lbfgs_trace_get(self) -> np.ndarray:
        n = int(self.lib.pinn_engine_lbfgs_trace_rows(self.h))
        out = np.empty((n, self.n_params), dtype=np.float32)
        _check(self.lib, self.lib.pinn_engine_lbfgs_trace_get(self.h, _ptr(out), n))
        return out

    def lbfgs_host_syncs(self) -> int:
        return int(self.lib.pinn_engine_lbfgs_host_syncs(self.h))

    # ---- NCCL
    def init_nccl(self, unique_id: bytes, rank: int, world: int):
        buf = (C.c_uint8 * 128).from_buffer_copy(unique_id)
        _check(self.lib, self.lib.pinn_engine_init_nccl(self.h, buf, rank, world))

    @staticmethod
    def nccl_unique_id() -> bytes:
        lib = load_library()
        buf = (C.c_uint8 * 128)()
        _check(lib, lib.pinn_nccl_unique_id(buf))
        return bytes(buf)


def sample_lhs_device(n: int, lo: Sequence[float], hi: Sequence[float], seed: int, device: int = 0, out=None, col0: int = 0):
    """Latin-hypercube points on the device (pyDOE.lhs, software.py:553,562): torch CUDA tensor [n, d]."""
    import torch

    lib = load_library()
    d = len(lo)
    if out is None:
        out = torch.empty((n, d), dtype=torch.float32, device=f"cuda:{device}")
    lo_c, hi_c = (C.c_float * 3)(*([float(v) for v in lo] + [0.0] * (3 - d))), (C.c_float * 3)(*([float(v) for v in hi] + [0.0] * (3 - d)))
    st = torch.cuda.current_stream(device).cuda_stream
    _check(lib, lib.pinn_sample_lhs(device, C.c_void_p(st), seed & 0xFFFFFFFF, n, d, lo_c, hi_c, C.c_void_p(out.data_ptr()),
                                    out.stride(0), col0))
    return out


def sample_cdf2d_device(n: int, X: np.ndarray, Y: np.ndarray, F: np.ndarray, seed: int, device: int = 0):
    """colloc2D_set (software.py:87-136) on the device: n points from the cell distribution F on grid X, Y."""
    import torch

    lib = load_library()
    Fc = np.asarray(F, dtype=np.float64)[0:-1, 0:-1]
    cum = np.ascontiguousarray(np.hstack([0.0, np.cumsum(Fc.reshape(-1))]))
    out = torch.empty((n, 2), dtype=torch.float32, device=f"cuda:{device}")
    st = torch.cuda.current_stream(device).cuda_stream
    _check(lib, lib.pinn_sample_cdf2d(device, C.c_void_p(st), seed & 0xFFFFFFFF, n, cum.ctypes.data_as(C.c_void_p),
                                      Fc.shape[0], Fc.shape[1], float(X[0, 0]), float(Y[0, 0]), float(X[0, 1] - X[0, 0]),
                                      float(Y[1, 0] - Y[0, 0]), C.c_void_p(out.data_ptr()), 2))
    return out


def fma_peak_tflops(device: int = 0, variant: int = 0) -> float:
    lib = load_library()
    out = C.c_double()
    _check(lib, lib.pinn_fma_peak(device, variant, C.byref(out)))
    return out.value


def lbfgs_direction(g, S, Y, rho, cnt: int, head: int, device: int = 0) -> np.ndarray:
    """Test hook: d = -H g through the engine's vector-free two-loop recursion (csrc/lbfgs_dev.cu) for a history of
    `cnt` live pairs in the ring S, Y [m][n] (newest at slot head - 1), rho[i] = 1 / (s_i . y_i)."""
    lib = load_library()
    g = np.ascontiguousarray(g, dtype=np.float32)
    S = np.ascontiguousarray(S, dtype=np.float32)
    Y = np.ascontiguousarray(Y, dtype=np.float32)
    rho = np.ascontiguousarray(rho, dtype=np.float64)
    m, n = S.shape
    out = np.empty(n, dtype=np.float32)
    lib.pinn_lbfgs_direction_test.argtypes = [C.c_int, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p,
                                              C.c_void_p, C.c_void_p]
    _check(lib, lib.pinn_lbfgs_direction_test(device, n, m, int(cnt), int(head), _ptr(g), _ptr(S), _ptr(Y), _ptr(rho), _ptr(out)))
    return out


def shard_range(n: int, rank: int, world: int):
    """Contiguous equal shards: [begin, end) of rank."""
    base, rem = divmod(n, world)
    b = rank * base + min(rank, rem)
    return b, b + base + (1 if rank < rem else 0)
