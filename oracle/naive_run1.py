"""Second, independent restatement of Run1's feature assembly and export transform in numpy
(float32 element ops in the Fortran evaluation order).  TEST INFRASTRUCTURE ONLY; PARITY
UNPINNED (see oracle/qc_oracle.h).  It exists so that oracle/qc_oracle.c (C), this file (numpy)
and the CUDA kernels are three separately written implementations that must agree bit for bit.

Follows /root/reference/OH_GridComp/OH_GridCompMod.F90:
  :1247-1257 PL_MOD, TV_MOD, NDWET_MOD        :1444-1466 latarr, stratO3, gridBoxThickness, aod
  :1468-1478 six vertical SUM(...,3)           :401-466   noon SZA (+ JulianDay :1905-1936)
  :275-301   level slab                         :308-345   pack order        :1569-1595 export
"""
from __future__ import annotations

import math

import numpy as np

f32 = np.float32


def leap_year(ny: int) -> bool:
    if ny >= 0:
        if ny % 100 == 0 and ny % 400 == 0:
            return True
        if ny % 4 == 0 and ny % 100 != 0:
            return True
    return False


def julian_day(nymd: int) -> int:
    days = [31, 28, 31, 30, 31, 30, 31, 31, 30, 31, 30, 31]
    ny, mm, dd = nymd // 10000, (nymd % 10000) // 100, nymd % 100
    ds = dd
    for m in range(1, mm):
        ds += 29 if (m == 2 and leap_year(ny)) else days[m - 1]
    return ds


def forward_sums(x: np.ndarray):
    """(UP, DN) with UP[k] = SUM(x[0..k]) and DN[k] = SUM(x[k..km-1]), each a fresh float32
    left-to-right chain as the Fortran SUM(...,3) evaluates it (:1468-1478)."""
    km = x.shape[0]
    up = np.empty_like(x)
    dn = np.empty_like(x)
    for k in range(km):
        s = np.zeros(x.shape[1:], f32)
        for kk in range(0, k + 1):
            s = (s + x[kk]).astype(f32)
        up[k] = s
        s = np.zeros(x.shape[1:], f32)
        for kk in range(k, km):
            s = (s + x[kk]).astype(f32)
        dn[k] = s
    return up, dn


def features(fields: dict, consts: dict, nymd: int, sza_deg: np.ndarray):
    """All 27 features as [km][ncol] (3-D) or [ncol] (2-D) float32 arrays, feature order of
    OH_BOOST_INPUT_DATA (:82-114).  `sza_deg` is supplied by the caller (libm-dependent)."""
    ple, zle = fields["PLE"], fields["ZLE"]
    pl = ((ple[:-1] + ple[1:]) * f32(0.5)).astype(f32)
    thick = (zle[:-1] - zle[1:]).astype(f32)
    s = (fields["BCSCACOEF"] + fields["OCSCACOEF"]).astype(f32)
    for sp in ("BR", "DU", "SU", "SS", "NI"):
        s = (s + fields[sp + "SCACOEF"]).astype(f32)
    aod = (thick.astype(np.float64) * s.astype(np.float64)).astype(f32)  # REAL*8 thickness * REAL
    wup, wdn = forward_sums(fields["TAUCLW"])
    iup, idn = forward_sums(fields["TAUCLI"])
    aup, adn = forward_sums(aod)
    lat = (fields["LATS"] * consts["RADIANS_TO_DEGREES"]).astype(f32)
    so3 = (fields["oh_GMITO3"] - fields["oh_GMITTO3"]).astype(f32)
    return [lat, pl, fields["T"], fields["oh_NO2"], fields["oh_O3"], fields["CH4"], fields["CO"], fields["oh_ISOP"],
            fields["oh_ACET"], fields["oh_C2H6"], fields["oh_C3H8"], fields["oh_PRPE"], fields["oh_ALK4"],
            fields["oh_MP"], fields["oh_H2O2"], wdn, idn, iup, wup, fields["FCLD"], fields["Q"], so3,
            fields["oh_ALBUV"], aup, adn, fields["oh_CH2O"], sza_deg]  # fmt: skip


def level_slab(pl_mod, tropp, compute_once_per_day, tropp_min):
    cmp = np.full_like(tropp, tropp_min) if compute_once_per_day else tropp
    ksub = int((pl_mod > cmp[None, :]).sum(axis=0).max())
    return pl_mod.shape[0] - ksub + 1  # k1, 1-based


def pack(feats, k1: int):
    km, ncol = feats[2].shape
    rows = (km - k1 + 1) * ncol
    x = np.empty((rows, 27), f32)
    for j, a in enumerate(feats):
        if a.ndim == 1:
            x[:, j] = np.tile(a, km - k1 + 1)
        else:
            x[:, j] = a[k1 - 1 :].reshape(-1)
    x[:, 1] = (x[:, 1] / f32(100.0)).astype(f32)
    return x


def export(pred, k1, fields, consts, ohscale, mod=None):
    mod = mod or fields
    km, ncol = fields["T"].shape
    pl = ((mod["PLE"][:-1] + mod["PLE"][1:]) * f32(0.5)).astype(f32)
    tv = (mod["T"] * (f32(1.0) + mod["Q"] / consts["EPSILON"]) / (f32(1.0) + mod["Q"])).astype(f32)
    ndwet = ((consts["AVOGAD"] * pl) / (consts["RUNIV"] * tv)).astype(f32)
    oh_ml = np.zeros((km, ncol), f32)
    oh_ml[k1 - 1 :] = np.array([math.pow(10.0, float(p)) for p in pred], np.float64).astype(f32).reshape(-1, ncol)
    oh_ml = (oh_ml * f32(ohscale)).astype(f32)
    oh = np.where(pl > mod["TROPP"][None, :], oh_ml, fields["oh_OH"]).astype(f32)
    return ((oh * ndwet).astype(f32) * f32(1.0e-6)).astype(f32), oh_ml, ndwet
