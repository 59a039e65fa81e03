"""Frames, body-fixed orientation models and the ephemeris container (``Almanac``).

In the reference all of this comes from anise 0.10.2 (`Frame`, `Almanac::{frame_info,
transform, rotate}`) reading `pck08.pca` / `de440s.bsp`, neither of which exists in the
reference tree (git-LFS stubs).  This module is therefore an explicit, documented model:

* ``Frame`` carries what the hot path reads from an anise Frame: ``mu_km3_s2()``
  (orbital.rs:86-90), ``mean_equatorial_radius_km()`` (gravity_field.rs:195-200), and an
  orientation (``Rotation``) when the frame is body-fixed.
* ``Rotation`` is the IAU pole / prime-meridian model declared in ``include/nyxb.h``.
* ``Almanac`` holds piecewise-Chebyshev ephemerides of celestial bodies relative to the
  integration centre (the data anise would interpolate from an SPK file).

Epochs are integer nanoseconds past J2000 (hifitime `Epoch` is an integer-ns type).
"""
from __future__ import annotations

from dataclasses import dataclass, field, replace
from typing import Dict, List, Optional

import numpy as np

NS_PER_S = 1_000_000_000
NS_PER_DAY = 86_400 * NS_PER_S


@dataclass(frozen=True)
class Rotation:
    """Orientation of a body-fixed frame: angles in degrees, T in Julian centuries, d in days."""

    ra0_deg: float = 0.0
    ra1_deg_cy: float = 0.0
    dec0_deg: float = 90.0
    dec1_deg_cy: float = 0.0
    w0_deg: float = 0.0
    w1_deg_day: float = 0.0
    kind: int = 1  # 0 = identity, 1 = IAU model

    @classmethod
    def identity(cls) -> "Rotation":
        return cls(kind=0)


# IAU WGCCRE constants as distributed in NAIF pck00008.tpc (the reference's pck08.pca).
IAU_EARTH_ROTATION = Rotation(0.0, -0.641, 90.0, -0.557, 190.147, 360.9856235)
# Mean terms only (no libration series): documented simplification for the Moon.
IAU_MOON_ROTATION = Rotation(269.9949, 0.0031, 66.5392, 0.0130, 38.3213, 13.17635815)


@dataclass(frozen=True)
class Frame:
    """Subset of anise's ``Frame`` used on the propagation path."""

    name: str
    ephemeris_id: int
    mu: Optional[float] = None  # km^3/s^2
    radius_km: Optional[float] = None  # mean equatorial radius
    rotation: Optional[Rotation] = None  # None: inertial (J2000 axes)
    polar_radius_km: Optional[float] = None  # ellipsoid shape for geodetic coordinates (None: sphere)

    def mu_km3_s2(self) -> float:
        if self.mu is None:
            raise ValueError(f"frame {self.name}: gravitational parameter not set")  # AstroPhysicsError
        return self.mu

    def mean_equatorial_radius_km(self) -> float:
        if self.radius_km is None:
            raise ValueError(f"frame {self.name}: shape not set")
        return self.radius_km

    def with_mu_km3_s2(self, mu: float) -> "Frame":
        return replace(self, mu=mu)

    def with_radius_km(self, r: float) -> "Frame":
        return replace(self, radius_km=r)


# NAIF ids
SUN, MOON, EARTH, JUPITER_BARYCENTER = 10, 301, 399, 5

# GM values: DE440 (pck08.pca's Earth GM reproduces orbitaldyn.rs:112-119 bit-exactly).
SUN_J2000 = Frame("Sun J2000", SUN, 132712440041.27942, 696000.0)
EARTH_J2000 = Frame("Earth J2000", EARTH, 398600.435436096, 6378.14)
MOON_J2000 = Frame("Moon J2000", MOON, 4902.800066163796, 1737.4)
JUPITER_BARYCENTER_J2000 = Frame("Jupiter Barycenter J2000", JUPITER_BARYCENTER, 126712764.09999998, 71492.0)
IAU_EARTH_FRAME = Frame("IAU Earth", EARTH, 398600.435436096, 6378.14, IAU_EARTH_ROTATION, 6356.75)  # pck00008 radii
IAU_MOON_FRAME = Frame("IAU Moon", MOON, 4902.800066163796, 1737.4, IAU_MOON_ROTATION)

_FRAMES = {f.ephemeris_id: f for f in (SUN_J2000, EARTH_J2000, MOON_J2000, JUPITER_BARYCENTER_J2000)}

# GMAT constants used by the reference tests (tests/propagation/mod.rs:1-3)
GMAT_EARTH_GM = 398_600.441_5
GMAT_SUN_GM = 132_712_440_017.99
GMAT_MOON_GM = 4_902.800_582_147_8


@dataclass
class BodyEphemeris:
    """Piecewise Chebyshev position of one body w.r.t. the almanac centre (J2000 axes)."""

    frame: Frame
    t0_ns: int
    interval_ns: int
    coeffs: np.ndarray  # [n_intervals, 3, n_coeffs] float64, C-contiguous

    @property
    def n_intervals(self) -> int:
        return self.coeffs.shape[0]

    @property
    def n_coeffs(self) -> int:
        return self.coeffs.shape[2]

    def position(self, t_ns: int) -> np.ndarray:
        """Reference (numpy) evaluation; same Clenshaw recurrence as the kernels."""
        idx, off = divmod(int(t_ns) - self.t0_ns, self.interval_ns)
        if idx < 0 or idx >= self.n_intervals:
            raise ValueError("epoch outside ephemeris coverage")
        tau = 2.0 * (off / self.interval_ns) - 1.0
        out = np.empty(3)
        for ax in range(3):
            c = self.coeffs[idx, ax]
            b1 = b2 = 0.0
            for k in range(self.n_coeffs - 1, 0, -1):
                b1, b2 = (2.0 * tau * b1 - b2) + c[k], b1
            out[ax] = (tau * b1 - b2) + c[0]
        return out


@dataclass
class Almanac:
    """Ephemerides + frame constants for one integration centre (stand-in for anise's Almanac)."""

    center: Frame
    bodies: List[BodyEphemeris] = field(default_factory=list)
    frames: Dict[int, Frame] = field(default_factory=lambda: dict(_FRAMES))

    def frame_info(self, frame: Frame) -> Frame:
        """`Almanac::frame_info`: fill in mu / shape for an id (orbital.rs:223-227)."""
        known = self.frames.get(frame.ephemeris_id)
        if known is None:
            raise KeyError(f"planetary data for {frame.name} not loaded")
        return replace(frame, mu=frame.mu if frame.mu is not None else known.mu,
                       radius_km=frame.radius_km if frame.radius_km is not None else known.radius_km)

    def body_index(self, ephemeris_id: int) -> int:
        for i, b in enumerate(self.bodies):
            if b.frame.ephemeris_id == ephemeris_id:
                return i
        raise KeyError(f"no ephemeris loaded for body {ephemeris_id}")

    def has_body(self, ephemeris_id: int) -> bool:
        return any(b.frame.ephemeris_id == ephemeris_id for b in self.bodies)

    @classmethod
    def synthetic(cls, center: Frame = EARTH_J2000, t0_ns: int = 0, span_days: float = 40.0,
                  bodies=(SUN, MOON), pad_days: float = 2.0) -> "Almanac":
        """Build Chebyshev ephemerides from the analytic series in ``nyx_b200.ephem``."""
        from . import ephem

        alm = cls(center=center)
        start = int(t0_ns) - int(pad_days * NS_PER_DAY)
        for bid in bodies:
            if bid == center.ephemeris_id:
                continue
            alm.bodies.append(ephem.chebyshev_ephemeris(bid, center.ephemeris_id, start, span_days + 2 * pad_days,
                                                       frame=_FRAMES[bid]))
        return alm
