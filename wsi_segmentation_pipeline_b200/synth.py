"""Deterministic synthetic H&E-statistics slide generator (host side, numpy).

SURVEY.md §8(d) asks for seeded, reproducible slides whose bytes are identical on the CPU
(oracle / golden fixtures / CPU baseline) and on the device (bench, GPU tests) without ever
materialising a 24 GB raster on the host.  Everything here is therefore *integer only* and
*counter based*: a pixel's bytes are a pure function of ``(seed, y, x)`` so any row band can
be produced independently.  The CUDA twin is ``wsi_synth_slide`` in ``csrc/synth.cuh``;
``tests/test_synth.py`` pins the two against each other and against a committed checksum.

Statistics (not a stain simulator, just H&E-like first/second order statistics):
  * background ~ (240,240,240) +- 4
  * tissue = thresholded two-octave lattice value noise (~55 % area)
  * eosin optical density from low-frequency noise, haematoxylin from a jittered-grid
    nuclei point process (r = 3..6 px); colours through a 16x8 Beer-Lambert LUT built from
    the Ruifrok stain vectors H=(0.650,0.704,0.286), E=(0.072,0.990,0.105)
  * triangular pixel noise, clipped to u8
"""
from __future__ import annotations

import numpy as np

_U32 = np.uint32
NUC_CELL = 12            # nuclei jitter-grid pitch in pixels
TISSUE_THRESH = 30200    # on the 16-bit two-octave field; ~55 % tissue
N_E, N_H = 16, 8         # LUT levels


def _mix32(x: np.ndarray) -> np.ndarray:
    """lowbias32 finaliser on uint32 arrays (wrap-around arithmetic)."""
    x = x.astype(_U32, copy=True)
    x ^= x >> _U32(16)
    x *= _U32(0x7FEB352D)
    x ^= x >> _U32(15)
    x *= _U32(0x846CA68B)
    x ^= x >> _U32(16)
    return x


def _hash2(seed: int, a, b, salt: int) -> np.ndarray:
    """One-round lattice hash: mix32(a*K1 ^ b*K2 ^ (seed*K3 + salt))."""
    a = np.asarray(a).astype(_U32)
    b = np.asarray(b).astype(_U32)
    k = _U32((seed * 0xC2B2AE3D + salt * 0x27D4EB2F) & 0xFFFFFFFF)
    return _mix32((a * _U32(0x9E3779B1)) ^ (b * _U32(0x85EBCA77)) ^ k)


def _value_noise(seed: int, ys: np.ndarray, xs: np.ndarray, shift: int, salt: int) -> np.ndarray:
    """Bilinear lattice noise, cell = 2**shift px, result in [0, 255 << 8] as uint32 [len(ys), len(xs)]."""
    cell = 1 << shift
    gy, fy = (ys >> shift), (ys & (cell - 1)).astype(_U32)
    gx, fx = (xs >> shift), (xs & (cell - 1)).astype(_U32)
    gy2, gx2 = gy[:, None], gx[None, :]
    v00 = _hash2(seed, gx2, gy2, salt) & _U32(255)
    v10 = _hash2(seed, gx2 + 1, gy2, salt) & _U32(255)
    v01 = _hash2(seed, gx2, gy2 + 1, salt) & _U32(255)
    v11 = _hash2(seed, gx2 + 1, gy2 + 1, salt) & _U32(255)
    fx2, fy2 = fx[None, :], fy[:, None]
    c = _U32(cell)
    top = v00 * (c - fx2) + v10 * fx2          # <= 255 * cell
    bot = v01 * (c - fx2) + v11 * fx2
    val = top * (c - fy2) + bot * fy2          # <= 255 * cell^2  (cell <= 512 -> < 2^26)
    return (val >> _U32(2 * shift - 8)).astype(_U32)   # [0, 255<<8]


def synth_lut() -> np.ndarray:
    """Beer-Lambert colour table u8 [N_E, N_H, 3]; computed once on the host in float64 and
    shipped to the device as bytes (so no device transcendental can perturb parity)."""
    H = np.array([0.650, 0.704, 0.286])
    E = np.array([0.072, 0.990, 0.105])
    lut = np.zeros((N_E, N_H, 3), np.uint8)
    for e in range(N_E):
        e_od = 0.20 + 0.70 * e / (N_E - 1)
        for h in range(N_H):
            h_od = 0.0 if h == 0 else 0.55 + 0.15 * (h - 1)
            od = e_od * E * 0.55 + h_od * H
            lut[e, h] = np.clip(np.floor(240.0 * np.power(10.0, -od) + 0.5), 0, 255).astype(np.uint8)
    return lut


def _tissue_field(seed: int, ys: np.ndarray, xs: np.ndarray) -> np.ndarray:
    f1 = _value_noise(seed, ys, xs, 9, 11)     # 512 px cells
    f2 = _value_noise(seed, ys, xs, 7, 23)     # 128 px cells
    return (f1 * _U32(3) + f2) >> _U32(2)      # [0, 65280]


def synth_mask(ih: int, iw: int, seed: int = 1234, y0: int = 0, y1: int | None = None) -> np.ndarray:
    """Foreground (tissue) indicator u8 {0,1} [y1-y0, iw] — the generator's ground truth, used
    as the stand-in for the reference's cached ``find_nuclei`` PNG (utils/dataset.py:131-138)."""
    y1 = ih if y1 is None else y1
    ys = np.arange(y0, y1, dtype=np.int64)
    xs = np.arange(0, iw, dtype=np.int64)
    return (_tissue_field(seed, ys, xs) > _U32(TISSUE_THRESH)).astype(np.uint8)


def synth_slide(ih: int, iw: int, seed: int = 1234, y0: int = 0, y1: int | None = None) -> np.ndarray:
    """RGB raster rows [y0, y1) of the (ih x iw) synthetic slide, u8 [y1-y0, iw, 3]."""
    y1 = ih if y1 is None else y1
    ys = np.arange(y0, y1, dtype=np.int64)
    xs = np.arange(0, iw, dtype=np.int64)
    tissue = _tissue_field(seed, ys, xs) > _U32(TISSUE_THRESH)

    # per-pixel hash (noise bits)
    hp = _hash2(seed, xs[None, :], ys[:, None], 37)

    # eosin level from 64 px lattice noise + 2 bits of pixel dither
    ef = _value_noise(seed, ys, xs, 6, 41)                       # [0, 65280]
    e_lvl = np.minimum((ef >> _U32(12)) + ((hp >> _U32(28)) & _U32(3)), _U32(N_E - 1)).astype(np.int64)

    # nuclei: jittered grid, test the 3x3 neighbouring cells
    cy, cx = ys // NUC_CELL, xs // NUC_CELL
    h_lvl = np.zeros((len(ys), len(xs)), np.int64)
    for dy in (-1, 0, 1):
        for dx in (-1, 0, 1):
            gy, gx = (cy + dy)[:, None], (cx + dx)[None, :]
            hc = _hash2(seed, gx, gy, 53)
            present = (hc & _U32(7)) < _U32(5)
            ox = ((hc >> _U32(3)) % _U32(NUC_CELL)).astype(np.int64)
            oy = ((hc >> _U32(9)) % _U32(NUC_CELL)).astype(np.int64)
            r = 3 + ((hc >> _U32(15)) & _U32(3)).astype(np.int64)
            ddx = xs[None, :] - (gx * NUC_CELL + ox)
            ddy = ys[:, None] - (gy * NUC_CELL + oy)
            inside = present & (ddx * ddx + ddy * ddy <= r * r)
            lvl = 1 + ((hc >> _U32(17)) % _U32(N_H - 1)).astype(np.int64)
            h_lvl = np.where(inside, np.maximum(h_lvl, lvl), h_lvl)

    lut = synth_lut().astype(np.int64)
    base = lut[e_lvl, h_lvl]                                      # [h, w, 3]

    out = np.empty((len(ys), len(xs), 3), np.int64)
    for c in range(3):
        bits = (hp >> _U32(6 * c)) & _U32(63)
        tri = ((bits & _U32(7)) + (bits >> _U32(3))).astype(np.int64) - 7      # [-7, 7]
        bg = 240 + ((hp >> _U32(4 * c + 8)) % _U32(9)).astype(np.int64) - 4
        out[..., c] = np.where(tissue, base[..., c] + tri, bg)
    return np.clip(out, 0, 255).astype(np.uint8)


def checksum(a: np.ndarray) -> int:
    """Order-sensitive 64-bit checksum used by fixtures (FNV-1a over bytes, vectorised in blocks)."""
    b = np.ascontiguousarray(a).view(np.uint8).ravel().astype(np.uint64)
    idx = np.arange(1, b.size + 1, dtype=np.uint64)
    with np.errstate(over="ignore"):
        return int(((b + np.uint64(1)) * (idx * np.uint64(0x9E3779B97F4A7C15) | np.uint64(1))).sum(dtype=np.uint64))
