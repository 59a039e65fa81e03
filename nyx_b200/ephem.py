"""Analytic Sun / Moon / Jupiter positions and their Chebyshev tabulation.

DE440s (the reference's SPK, read through anise) is a git-LFS stub in the reference tree,
so the ephemerides fed to the engine are generated here from low-precision analytic series
(Montenbruck & Gill, *Satellite Orbits*, §3.3.2, for the Sun and the Moon; Standish's mean
Keplerian elements for Jupiter) and tabulated as piecewise Chebyshev polynomials — the same
representation an SPK type-2 segment uses.  CPU oracle and GPU kernels evaluate the SAME
tables, so parity does not depend on the accuracy of the series (few 1e-4 rad).
"""
from __future__ import annotations

import numpy as np

from .frames import (EARTH, JUPITER_BARYCENTER, MOON, NS_PER_DAY, NS_PER_S, SUN, BodyEphemeris, Frame)

_ARCSEC = np.pi / (180.0 * 3600.0)
_DEG = np.pi / 180.0
_EPS = 23.43929111 * _DEG  # obliquity of the ecliptic at J2000
_AU = 149_597_870.700


def _ecl_to_equ(x, y, z):
    ce, se = np.cos(_EPS), np.sin(_EPS)
    return np.stack([x, ce * y - se * z, se * y + ce * z], axis=-1)


def sun_wrt_earth(t_s):
    """Geocentric Sun position [km], J2000 equatorial axes; t_s = seconds past J2000."""
    T = np.asarray(t_s, dtype=np.float64) / (86400.0 * 36525.0)
    M = (357.5256 + 35999.049 * T) * _DEG
    lam = (282.9400 * _DEG) + M + (6892.0 * np.sin(M) + 72.0 * np.sin(2 * M)) * _ARCSEC
    r = (149.619 - 2.499 * np.cos(M) - 0.021 * np.cos(2 * M)) * 1e6
    return _ecl_to_equ(r * np.cos(lam), r * np.sin(lam), np.zeros_like(r))


def moon_wrt_earth(t_s):
    """Geocentric Moon position [km], J2000 equatorial axes (M&G eqs. 3.47-3.51)."""
    T = np.asarray(t_s, dtype=np.float64) / (86400.0 * 36525.0)
    L0 = (218.31617 + 481267.88088 * T - 1.3972 * T) * _DEG
    l = (134.96292 + 477198.86753 * T) * _DEG
    lp = (357.52543 + 35999.04944 * T) * _DEG
    F = (93.27283 + 483202.01873 * T) * _DEG
    D = (297.85027 + 445267.11135 * T) * _DEG
    dlam = (22640 * np.sin(l) + 769 * np.sin(2 * l) - 4586 * np.sin(l - 2 * D) + 2370 * np.sin(2 * D)
            - 668 * np.sin(lp) - 412 * np.sin(2 * F) - 212 * np.sin(2 * l - 2 * D) - 206 * np.sin(l + lp - 2 * D)
            + 192 * np.sin(l + 2 * D) - 165 * np.sin(lp - 2 * D) + 148 * np.sin(l - lp) - 125 * np.sin(D)
            - 110 * np.sin(l + lp) - 55 * np.sin(2 * F - 2 * D)) * _ARCSEC
    lam = L0 + dlam
    beta = (18520 * np.sin(F + dlam + (412 * np.sin(2 * F) + 541 * np.sin(lp)) * _ARCSEC)
            - 526 * np.sin(F - 2 * D) + 44 * np.sin(l + F - 2 * D) - 31 * np.sin(-l + F - 2 * D)
            - 25 * np.sin(-2 * l + F) - 23 * np.sin(lp + F - 2 * D) + 21 * np.sin(-l + F)
            + 11 * np.sin(-lp + F - 2 * D)) * _ARCSEC
    r = (385000 - 20905 * np.cos(l) - 3699 * np.cos(2 * D - l) - 2956 * np.cos(2 * D) - 570 * np.cos(2 * l)
         + 246 * np.cos(2 * l - 2 * D) - 205 * np.cos(lp - 2 * D) - 171 * np.cos(l + 2 * D)
         - 152 * np.cos(l + lp - 2 * D))
    cb = np.cos(beta)
    return _ecl_to_equ(r * np.cos(lam) * cb, r * np.sin(lam) * cb, r * np.sin(beta))


def jupiter_wrt_sun(t_s):
    """Heliocentric Jupiter-barycentre position [km] from mean Keplerian elements (Standish)."""
    T = np.asarray(t_s, dtype=np.float64) / (86400.0 * 36525.0)
    a = (5.20288700 - 0.00011607 * T) * _AU
    e = 0.04838624 - 0.00013253 * T
    inc = (1.30439695 - 0.00183714 * T) * _DEG
    L = (34.39644051 + 3034.74612775 * T) * _DEG
    varpi = (14.72847983 + 0.21252668 * T) * _DEG
    Om = (100.47390909 + 0.20469106 * T) * _DEG
    w = varpi - Om
    M = np.mod(L - varpi + np.pi, 2 * np.pi) - np.pi
    E = M + e * np.sin(M)
    for _ in range(8):
        E = E - (E - e * np.sin(E) - M) / (1.0 - e * np.cos(E))
    xp = a * (np.cos(E) - e)
    yp = a * np.sqrt(1.0 - e * e) * np.sin(E)
    cw, sw, cO, sO, ci, si = np.cos(w), np.sin(w), np.cos(Om), np.sin(Om), np.cos(inc), np.sin(inc)
    x = (cw * cO - sw * sO * ci) * xp + (-sw * cO - cw * sO * ci) * yp
    y = (cw * sO + sw * cO * ci) * xp + (-sw * sO + cw * cO * ci) * yp
    z = (sw * si) * xp + (cw * si) * yp
    return _ecl_to_equ(x, y, z)


def position(body: int, center: int, t_s):
    """Position of `body` relative to `center` (NAIF ids), km, J2000 axes."""

    def wrt_earth(b):
        if b == EARTH:
            return np.zeros(np.shape(t_s) + (3,))
        if b == SUN:
            return sun_wrt_earth(t_s)
        if b == MOON:
            return moon_wrt_earth(t_s)
        if b == JUPITER_BARYCENTER:
            return sun_wrt_earth(t_s) + jupiter_wrt_sun(t_s)
        raise KeyError(f"no analytic ephemeris for body {b}")

    return wrt_earth(body) - wrt_earth(center)


# (interval_days, n_coeffs) per body: chosen so the interpolation error is far below the series error
_TABULATION = {SUN: (16.0, 12), MOON: (4.0, 14), JUPITER_BARYCENTER: (32.0, 10), EARTH: (4.0, 14)}


def chebyshev_ephemeris(body: int, center: int, t0_ns: int, span_days: float, frame: Frame) -> BodyEphemeris:
    interval_days, nc = _TABULATION[body]
    if center == MOON or body == MOON:
        interval_days, nc = _TABULATION[MOON]
    interval_ns = int(interval_days * NS_PER_DAY)
    n_int = int(np.ceil(span_days * NS_PER_DAY / interval_ns))
    # Chebyshev-Gauss nodes and the discrete orthogonality sums
    j = np.arange(nc)
    x = np.cos(np.pi * (j + 0.5) / nc)  # nodes in (-1, 1)
    Tk = np.cos(np.outer(np.arange(nc), np.arccos(x)))  # [k, j]
    coeffs = np.empty((n_int, 3, nc))
    for i in range(n_int):
        start_s = (t0_ns + i * interval_ns) / NS_PER_S
        ts = start_s + 0.5 * (x + 1.0) * (interval_ns / NS_PER_S)
        f = position(body, center, ts)  # [j, 3]
        c = (2.0 / nc) * (Tk @ f)  # [k, 3]
        c[0] *= 0.5
        coeffs[i] = c.T
    return BodyEphemeris(frame=frame, t0_ns=int(t0_ns), interval_ns=interval_ns, coeffs=np.ascontiguousarray(coeffs))
