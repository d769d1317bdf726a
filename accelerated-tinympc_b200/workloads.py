"""Seeded synthetic batches for the BASELINE.json configs.

The generator is index-based (O(1) per element) so any shard [b0, b1) of a batch can be produced
independently on any rank:  u01(seed, idx) = (splitmix64(seed ^ splitmix64(idx)) >> 40) * 2^-24,
which is exactly representable in fp32.  Initial states are formed in float64 and rounded once to
float32, so the f32 and f64 solvers see identical inputs.
"""
from __future__ import annotations

import numpy as np

from . import problems

_M64 = np.uint64(0xFFFFFFFFFFFFFFFF)


def splitmix64(x):
    x = np.asarray(x, dtype=np.uint64)
    with np.errstate(over="ignore"):
        x = x + np.uint64(0x9E3779B97F4A7C15)
        x = (x ^ (x >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        x = (x ^ (x >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        return x ^ (x >> np.uint64(31))


def u01(seed: int, idx):
    h = splitmix64(np.uint64(seed) ^ splitmix64(idx))
    return (h >> np.uint64(40)).astype(np.float64) * (2.0 ** -24)


QUAD_SCALE = np.array([2, 2, 2, .2, .2, .2, .5, .5, .5, .5, .5, .5], dtype=np.float64)
QUAD_HOVER = np.array([0, 0, 2, 0, 0, 0, 0, 0, 0, 0, 0, 0], dtype=np.float64)


def _noise(seed, b0, b1, n):
    idx = (np.arange(b0, b1, dtype=np.uint64)[:, None] * np.uint64(n) + np.arange(n, dtype=np.uint64)[None, :])
    return 2.0 * u01(seed, idx) - 1.0


def quadrotor_hover_batch(b0: int, b1: int, mult: float = 0.25, N: int = 10, seed: int = 1234):
    """Config 2: random initial states around the hover set-point, one shared Xref.
    Returns (x0 float32 [B,12], Xref float32 [N,12])."""
    x0 = QUAD_HOVER[None, :] + mult * QUAD_SCALE[None, :] * _noise(seed, b0, b1, 12)
    xref = np.tile(QUAD_HOVER[None, :], (N, 1))
    return x0.astype(np.float32), xref.astype(np.float32)


def quadrotor_tracking_batch(b0: int, b1: int, N: int = 10, seed: int = 4321, mult: float = 0.1):
    """Config 3: per-instance reference windows k_b = b mod 290 of the y-axis-line table
    (examples/quadrotor_tracking.cpp:93,101).  Returns (x0 [B,12], Xref [B,N,12]) float32."""
    table = problems.quadrotor_trajectory()                    # (12, 301)
    kb = np.arange(b0, b1, dtype=np.int64) % 290
    win = kb[:, None] + np.arange(N)[None, :]                   # [B, N]
    xref = table.T[win]                                         # [B, N, 12]
    x0 = table.T[kb] + mult * QUAD_SCALE[None, :] * _noise(seed, b0, b1, 12)
    return x0.astype(np.float32), np.ascontiguousarray(xref.astype(np.float32))


CART_SCALE = np.array([0.5, 0.5, 0.2, 0.5], dtype=np.float64)


def cartpole_batch(b0: int, b1: int, N: int = 10, seed: int = 777):
    """Config 4: x0 = scale * (2 u01 - 1), Xref = 0."""
    x0 = CART_SCALE[None, :] * _noise(seed, b0, b1, 4)
    return x0.astype(np.float32), np.zeros((N, 4), dtype=np.float32)


def random_system_batch(b0: int, b1: int, N: int = 50, nx: int = 32, amp: float = 1.0, seed: int = 555):
    """Config 5 (problems.random_system, 32/8/50): x0 = amp * (2 u01 - 1), Xref = 0.  amp = 1.0 gives a cold
    solve with mean 71 iterations (71 % converge) and a warm-started re-solve with mean 27 (95 % converge)."""
    x0 = amp * _noise(seed, b0, b1, nx)
    return x0.astype(np.float32), np.zeros((N, nx), dtype=np.float32)


def perturb_x0(x0, b0: int, rel: float = 0.01, seed: int = 556):
    """Config 5's second solve: every component of x0 moved by rel * U(-1, 1) of itself (formed in float64)."""
    x0 = np.asarray(x0)
    n = x0.shape[1]
    return (x0.astype(np.float64) * (1.0 + rel * _noise(seed, b0, b0 + x0.shape[0], n))).astype(np.float32)
