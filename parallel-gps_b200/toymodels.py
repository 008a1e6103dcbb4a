"""Deterministic test signals and the noisy-observation helper of the reference's toy experiments
(pssgp/toymodels/data_funcs.py; BASELINE configs[0] is ``obs_noise(sinu(linspace(0, 4, 1000)), 0.1, seed=0)``).
Pinned bit-for-bit to vectors generated from the reference's own module (tests/golden/toy_sinusoid_n1000.npz)."""
import math

import numpy as np

__all__ = ["sinu", "comp_sinu", "rect", "obs_noise"]


def sinu(t):
    """sin(pi t) + sin(2 pi t) + cos(3 pi t)   (data_funcs.py:10-23)."""
    t = np.asarray(t)
    return np.sin(np.pi * t) + np.sin(2 * np.pi * t) + np.cos(3 * np.pi * t)


def comp_sinu(t):
    """sin^2(7 pi cos(2 pi t^2)) / (cos(5 pi t) + 2)   (data_funcs.py:26-43)."""
    t = np.asarray(t)
    return np.sin(7 * np.pi * np.cos(2 * np.pi * (t ** 2))) ** 2 / (np.cos(5 * np.pi * t) + 2)


def rect(t):
    """Piecewise-constant signal with levels 0, 1, 0, 0.6, 0, 0.4 on [0, 1/6), [1/6, 1/3), ... of the rescaled time axis
    (data_funcs.py:46-74)."""
    t = np.asarray(t)
    tau = (t - np.min(t)) / (np.max(t) - np.min(t))
    edges = np.linspace(1 / 6, 5 / 6, 5)
    levels = np.array([0.0, 1.0, 0.0, 0.6, 0.0, 0.4])
    return levels[np.searchsorted(edges, tau, side="right")]


def obs_noise(x, r, seed=None):
    """x + sqrt(r) * eps with eps ~ N(x, r) — the reference draws the noise with MEAN x (data_funcs.py:97), so the
    observations are x (1 + sqrt(r)) + r z; kept, because the published toy configuration is defined by it."""
    x = np.asarray(x)
    rng = np.random.RandomState(seed)
    return x + np.sqrt(r) * rng.normal(x, math.sqrt(r), (x.shape[0],)).astype(x.dtype)
