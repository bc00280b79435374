"""CPU oracle for SURVEY.md §8 row f2: blur / sinc kernel synthesis.  TEST INFRASTRUCTURE ONLY.

numpy (float64) restatement of the kernel generators the reference runs inside its dataset
workers — traiNNer/data/degradations.py:22-212 (Gaussian family), :472-507 (circular low-pass sinc)
— with the random draws made explicit, plus the zero padding and fp32 cast of
traiNNer/data/realesrgan_dataset.py:171-172, :200-211.  Pinned by oracle/make_goldens.py against the
imported reference functions (bit-identical: same numpy / scipy calls).

A kernel is described by 8 numbers:  [type, ksize, sig_x, sig_y, theta, beta, omega_c, pad_to]
with type 0 iso, 1 aniso, 2 generalized_iso, 3 generalized_aniso, 4 plateau_iso, 5 plateau_aniso,
6 sinc, 7 pulse.
"""

from __future__ import annotations

import numpy as np
from scipy import special

TYPES = ("iso", "aniso", "generalized_iso", "generalized_aniso", "plateau_iso", "plateau_aniso", "sinc", "pulse")


def _grid(k: int) -> np.ndarray:
    # degradations.py:40-59 (mesh_grid): axis -k//2+1 .. k//2, x along columns
    ax = np.arange(-k // 2 + 1.0, k // 2 + 1.0)
    xx, yy = np.meshgrid(ax, ax)
    return np.hstack((xx.reshape((k * k, 1)), yy.reshape(k * k, 1))).reshape(k, k, 2)


def _sigma_matrix(kind: int, sx: float, sy: float, theta: float) -> np.ndarray:
    if kind in (0, 2, 4):  # isotropic: only sig_x is used (:120-121)
        return np.array([[sx**2, 0], [0, sx**2]])
    d = np.array([[sx**2, 0], [0, sy**2]])  # :33-37
    u = np.array([[np.cos(theta), -np.sin(theta)], [np.sin(theta), np.cos(theta)]])
    return np.dot(u, np.dot(d, u.T))


def gaussian_family(kind: int, k: int, sx: float, sy: float, theta: float, beta: float) -> np.ndarray:
    """bivariate_Gaussian (:96-128), bivariate_generalized_Gaussian (:131-169), bivariate_plateau (:172-212),
    followed by the second normalisation of the random_* wrappers (:258, :313, :369)."""
    grid = _grid(k)
    inv = np.linalg.inv(_sigma_matrix(kind, sx, sy, theta))
    q = np.sum(np.dot(grid, inv) * grid, 2)
    if kind in (0, 1):
        ker = np.exp(-0.5 * q)
    elif kind in (2, 3):
        ker = np.exp(-0.5 * np.power(q, beta))
    else:
        ker = np.reciprocal(np.power(q, beta) + 1)
    ker = ker / np.sum(ker)
    return ker / np.sum(ker)


def circular_lowpass(cutoff: float, k: int) -> np.ndarray:
    """degradations.py:472-507 without the padding."""
    with np.errstate(divide="ignore", invalid="ignore"):
        ker = np.fromfunction(
            lambda x, y: cutoff
            * special.j1(cutoff * np.sqrt((x - (k - 1) / 2) ** 2 + (y - (k - 1) / 2) ** 2))
            / (2 * np.pi * np.sqrt((x - (k - 1) / 2) ** 2 + (y - (k - 1) / 2) ** 2)),
            [k, k],
        )
    ker[(k - 1) // 2, (k - 1) // 2] = cutoff**2 / (4 * np.pi)
    return ker / np.sum(ker)


def synthesize(params: np.ndarray) -> np.ndarray:
    """(B,8) float64 parameter table -> (B,21,21) float32 kernels, zero-padded as the dataset does."""
    out = np.zeros((len(params), 21, 21), dtype=np.float32)
    for b, (kind, k, sx, sy, theta, beta, wc, _pad) in enumerate(params):
        kind, k = int(kind), int(k)
        if kind == 7:
            out[b, 10, 10] = 1.0
            continue
        ker = circular_lowpass(wc, k) if kind == 6 else gaussian_family(kind, k, sx, sy, theta, beta)
        p = (21 - k) // 2
        out[b] = np.pad(ker, ((p, p), (p, p))).astype(np.float32)
    return out
