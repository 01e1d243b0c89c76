"""Synthetic workloads of BASELINE.json `configs`.

Weights: truncated-normal(-2,2)*sqrt(2/(in+out)) for W and b (init_MLP,
software.py:142-154) from a numpy RandomState(1234) stream (the JAX threefry
stream of the reference cannot be reproduced -- SURVEY.md section 8c).
Points: uniform in the domain, seed 1234 (software.py:685).
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Dict, List, Tuple

import numpy as np

from .engine import NetworkSpec
from .equation import CompiledEquation, compile_equation


@dataclass
class Workload:
    name: str
    description: str
    net: NetworkSpec
    expr: str
    n_col: int
    n_bd: List[int]  # points per boundary group
    lw: float = 1.0

    @property
    def eq(self) -> CompiledEquation:
        return compile_equation(self.expr, d_in=self.net.d_in)

    def flops_per_point(self) -> Dict[str, float]:
        """Algorithmic FLOPs per collocation point per train step: 2K(2 M1 + M2)
        (SURVEY.md section 8d); boundary points the same with K=1."""
        eq_full = compile_equation(self.expr, d_in=self.net.d_in, combine_second=False)
        F, W, L = self.net.n_feat, self.net.width, self.net.n_hidden
        m1 = F * W + (L - 1) * W * W + W
        m2 = (L - 1) * W * W + W
        # K = channels of the SURVEY's accounting (one per derivative); K_exec = channels the kernel
        # propagates (Laplacian-type operators need one combined second-order channel)
        return dict(col=2.0 * eq_full.K * (2 * m1 + m2), bc=2.0 * (2 * m1 + m2), K=eq_full.K,
                    K_exec=self.eq.K, col_exec=2.0 * self.eq.K * (2 * m1 + m2))


def truncated_normal(rng: np.random.RandomState, shape, lo=-2.0, hi=2.0) -> np.ndarray:
    from scipy.special import erf, erfinv

    a, b = erf(lo / math.sqrt(2)), erf(hi / math.sqrt(2))
    u = rng.uniform(a, b, size=shape)
    return np.clip(math.sqrt(2) * erfinv(u), lo, hi)


def init_params(net: NetworkSpec, seed: int = 1234) -> np.ndarray:
    """Flat fp32 parameter vector in ravel_pytree order (software.py:142-154, 466)."""
    rng = np.random.RandomState(seed)
    out = []
    lw = net.layer_widths
    for i, o in zip(lw[:-1], lw[1:]):
        std = math.sqrt(2.0 / (i + o))
        out.append((truncated_normal(rng, (i, o)) * std).reshape(-1))
        out.append((truncated_normal(rng, (o,)) * std).reshape(-1))
    return np.concatenate(out).astype(np.float32)


def unflatten(net: NetworkSpec, flat: np.ndarray):
    """flat -> [[W, b], ...] (numpy views)."""
    out, o = [], 0
    lw = net.layer_widths
    for i, j in zip(lw[:-1], lw[1:]):
        W = flat[o:o + i * j].reshape(i, j)
        o += i * j
        b = flat[o:o + j]
        o += j
        out.append([W, b])
    return out


def box_boundaries(lb, ub, n_per: int, rng: np.random.RandomState, dims=None) -> List[np.ndarray]:
    """One group per face of the box (lower then upper, per dim)."""
    d = len(lb)
    groups = []
    for k in (dims if dims is not None else range(d)):
        for side in (lb[k], ub[k]):
            p = rng.uniform(size=(n_per, d)) * (np.asarray(ub) - np.asarray(lb)) + np.asarray(lb)
            p[:, k] = side
            groups.append(p.astype(np.float32))
    return groups


def make_workload(name: str, n_col: int = None) -> Workload:
    if name == "R0":
        net = NetworkSpec(6, 60, [0.1, 0.0], [1.0, 1.0], feature_map="polar", d_in=2)
        return Workload("R0", "reference smoke: polar Laplace, 6x60 tanh, 3-feature map", net,
                        "u_rr + 1/r*u_r + 1/(r**2)*u_tt", n_col or 5200, [100, 100], lw=0.05)
    if name == "C1":
        net = NetworkSpec(3, 20, [0.0], [1.0], feature_map="affine", d_in=1)
        return Workload("C1", "1D Poisson u''=-2 on [0,1], 3x20 tanh", net, "u_xx + 2", n_col or 1000, [1, 1])
    if name == "C2":
        net = NetworkSpec(4, 64, [0.0, 0.0], [1.0, 1.0], feature_map="affine", d_in=2)
        return Workload("C2", "2D Poisson on the unit square, 4x64 tanh MLP, 1M collocation + 40k boundary points",
                        net, "u_xx + u_yy + 2*y*(1-y) + 2*x*(1-x)", n_col or 1_000_000, [10_000] * 4)
    if name == "C3":
        net = NetworkSpec(8, 50, [-1.0, 0.0], [1.0, 1.0], feature_map="affine", d_in=2)
        return Workload("C3", "1D viscous Burgers (x,t), 8x50 tanh", net, "u_t + u*u_x - 0.003183*u_xx",
                        n_col or 4_000_000, [10_000] * 3)
    if name == "C4":
        net = NetworkSpec(6, 128, [0.0, 0.0], [1.0, 1.0], scl=8.0, act_first=1, act_hidden=1,
                          feature_map="affine", d_in=2)
        k2 = (8 * math.pi) ** 2
        return Workload("C4", "2D Helmholtz high-frequency source, 6x128 sin MLP", net,
                        f"u_xx + u_yy + {k2:.6f}*u + {k2:.6f}*sin(8*pi*x)*sin(8*pi*y)", n_col or 16_000_000,
                        [10_000] * 4)
    if name == "C5":
        net = NetworkSpec(5, 256, [0.0, 0.0, 0.0], [1.0, 1.0, 1.0], feature_map="affine", d_in=3)
        return Workload("C5", "2D heat equation (x,y,t), 5x256 tanh", net, "u_t - 0.1*(u_xx + u_yy)",
                        n_col or 8_000_000, [10_000] * 5)
    raise KeyError(name)


def make_points(wl: Workload, seed: int = 1234, rank: int = 0) -> Tuple[np.ndarray, List[np.ndarray], List[np.ndarray]]:
    """(x_col, x_bd[], u_bd[]) fp32, uniform in the domain; rank offsets the seed
    (each rank draws its own shard)."""
    rng = np.random.RandomState(seed + 7919 * rank)
    lb, ub = np.asarray(wl.net.lb, dtype=np.float64), np.asarray(wl.net.ub, dtype=np.float64)
    x_col = (rng.uniform(size=(wl.n_col, wl.net.d_in)) * (ub - lb) + lb).astype(np.float32)
    d = wl.net.d_in
    x_bd: List[np.ndarray] = []
    if wl.name == "C1":
        x_bd = [np.array([[0.0]], np.float32), np.array([[1.0]], np.float32)]
    elif wl.name == "R0":
        for side in (lb[0], ub[0]):
            p = rng.uniform(size=(wl.n_bd[0], d)) * (ub - lb) + lb
            p[:, 0] = side
            x_bd.append(p.astype(np.float32))
    elif wl.name == "C3":
        p = rng.uniform(size=(wl.n_bd[0], d)) * (ub - lb) + lb
        p[:, 1] = lb[1]
        x_bd.append(p.astype(np.float32))
        for side in (lb[0], ub[0]):
            p = rng.uniform(size=(wl.n_bd[0], d)) * (ub - lb) + lb
            p[:, 0] = side
            x_bd.append(p.astype(np.float32))
    elif wl.name == "C5":
        p = rng.uniform(size=(wl.n_bd[0], d)) * (ub - lb) + lb
        p[:, 2] = lb[2]
        x_bd.append(p.astype(np.float32))
        x_bd += box_boundaries(lb, ub, wl.n_bd[0], rng, dims=(0, 1))
    else:
        x_bd = box_boundaries(lb, ub, wl.n_bd[0], rng)
    u_bd = []
    for i, p in enumerate(x_bd):
        if wl.name == "R0":
            u_bd.append(np.full(len(p), 1.0 if i == 0 else 0.0, np.float32))
        elif wl.name == "C3" and i == 0:
            u_bd.append((-np.sin(np.pi * p[:, 0])).astype(np.float32))
        elif wl.name == "C5" and i == 0:
            u_bd.append((np.sin(np.pi * p[:, 0]) * np.sin(np.pi * p[:, 1])).astype(np.float32))
        else:
            u_bd.append(np.zeros(len(p), np.float32))
    return x_col, x_bd, u_bd
