"""`Orbit` / `Spacecraft` state containers and hifitime-like time helpers.

Mirrors ``cosmic/spacecraft.rs:115-143`` (Spacecraft), the anise types it embeds
(`Orbit`, `Mass`, `SRPData`, `DragData`) and the State vector layout
``[x, y, z, vx, vy, vz, Cr, Cd, prop_mass]`` (cosmic/spacecraft.rs:449-473).
"""
from __future__ import annotations

from dataclasses import dataclass, field, replace
from typing import Iterable, Sequence

import numpy as np

from .frames import NS_PER_S, Frame


# --------------------------------------------------------------------------- time (hifitime semantics)
class Unit:
    """`hifitime::Unit`: ``x * Unit.Second`` -> integer-ns Duration, truncated toward zero."""

    class _U:
        def __init__(self, ns: int):
            self.ns = ns

        def __rmul__(self, q) -> int:
            if isinstance(q, (int, np.integer)):
                return int(q) * self.ns
            return int(float(q) * float(self.ns))  # `(q * factor) as i128`

        __mul__ = __rmul__

    Nanosecond = _U(1)
    Microsecond = _U(1_000)
    Millisecond = _U(1_000_000)
    Second = _U(NS_PER_S)
    Minute = _U(60 * NS_PER_S)
    Hour = _U(3600 * NS_PER_S)
    Day = _U(86400 * NS_PER_S)


def duration_to_seconds(total_ns: int) -> float:
    """`Duration::to_seconds` for |duration| < 1 century: whole seconds + subsec * 1e-9."""
    ns_per_century = 3_155_760_000 * NS_PER_S
    cent, nanos = divmod(int(total_ns), ns_per_century)
    sec, sub = divmod(nanos, NS_PER_S)
    if cent == 0:
        return float(sec) + float(sub) * 1e-9
    return float(cent) * 3_155_760_000.0 + float(sec) + float(sub) * 1e-9


# TAI - UTC (s) from the given UTC date on (IERS Bulletin C; public data, what hifitime's leap-second table holds)
_LEAP_SECONDS = (("1972-01-01", 10), ("1972-07-01", 11), ("1973-01-01", 12), ("1974-01-01", 13), ("1975-01-01", 14),
                 ("1976-01-01", 15), ("1977-01-01", 16), ("1978-01-01", 17), ("1979-01-01", 18), ("1980-01-01", 19),
                 ("1981-07-01", 20), ("1982-07-01", 21), ("1983-07-01", 22), ("1985-07-01", 23), ("1988-01-01", 24),
                 ("1990-01-01", 25), ("1991-01-01", 26), ("1992-07-01", 27), ("1993-07-01", 28), ("1994-07-01", 29),
                 ("1996-01-01", 30), ("1997-07-01", 31), ("1999-01-01", 32), ("2006-01-01", 33), ("2009-01-01", 34),
                 ("2012-07-01", 35), ("2015-07-01", 36), ("2017-01-01", 37))
_TT_MINUS_TAI_NS = 32_184_000_000


def epochs_to_utc_iso(epoch_ns) -> np.ndarray:
    """ISO-8601 UTC strings of epochs given in integer ns past J2000 (2000-01-01T12:00:00 TDB), for the "Epoch (UTC)" column of
    the parquet exports (mc/results.rs:355-360: `epoch.to_time_scale(UTC).to_isoformat()`).  TDB is taken equal to TT (the
    periodic difference stays below 1.7 ms); inside a leap second the UTC label repeats the following second."""
    ep = np.asarray(epoch_ns, dtype=np.int64)
    j2000 = np.datetime64("2000-01-01T12:00:00", "ns")
    starts = np.array([(np.datetime64(d + "T00:00:00", "ns") - j2000).astype(np.int64) + dat * NS_PER_S + _TT_MINUS_TAI_NS
                       for d, dat in _LEAP_SECONDS], dtype=np.int64)   # TT instants at which each TAI-UTC value starts to apply
    dat = np.array([d for _, d in _LEAP_SECONDS], dtype=np.int64)
    k = np.clip(np.searchsorted(starts, ep, side="right") - 1, 0, None)
    utc = ep - _TT_MINUS_TAI_NS - dat[k] * NS_PER_S
    return np.datetime_as_string(j2000 + utc.astype("timedelta64[ns]"), unit="ns")


def utc_iso_to_epochs(iso) -> np.ndarray:
    """Inverse of `epochs_to_utc_iso` (`Epoch::from_gregorian_str` on the UTC ISO strings the parquet files hold): integer ns
    past J2000 TDB (= TT here)."""
    txt = [str(x).replace(" UTC", "").strip() for x in np.atleast_1d(iso)]
    utc = (np.array(txt, dtype="datetime64[ns]") - np.datetime64("2000-01-01T12:00:00", "ns")).astype(np.int64)
    j2000 = np.datetime64("2000-01-01T12:00:00", "ns")
    starts_utc = np.array([(np.datetime64(d + "T00:00:00", "ns") - j2000).astype(np.int64) for d, _ in _LEAP_SECONDS], dtype=np.int64)
    dat = np.array([d for _, d in _LEAP_SECONDS], dtype=np.int64)
    k = np.clip(np.searchsorted(starts_utc, utc, side="right") - 1, 0, None)
    return utc + dat[k] * NS_PER_S + _TT_MINUS_TAI_NS


# --------------------------------------------------------------------------- state
@dataclass(frozen=True)
class Orbit:
    """anise `Orbit` (CartesianState): km, km/s, epoch in integer ns past J2000."""

    x_km: float
    y_km: float
    z_km: float
    vx_km_s: float
    vy_km_s: float
    vz_km_s: float
    epoch_ns: int
    frame: Frame

    @classmethod
    def cartesian(cls, x, y, z, vx, vy, vz, epoch_ns, frame) -> "Orbit":
        return cls(float(x), float(y), float(z), float(vx), float(vy), float(vz), int(epoch_ns), frame)

    @classmethod
    def keplerian(cls, sma_km, ecc, inc_deg, raan_deg, aop_deg, ta_deg, epoch_ns, frame) -> "Orbit":
        """anise `Orbit::keplerian` (classical elements -> Cartesian)."""
        mu = frame.mu_km3_s2()
        inc, raan, aop, ta = np.radians([inc_deg, raan_deg, aop_deg, ta_deg])
        p = sma_km * (1.0 - ecc * ecc)
        r = p / (1.0 + ecc * np.cos(ta))
        rp = np.array([r * np.cos(ta), r * np.sin(ta), 0.0])
        vp = np.sqrt(mu / p) * np.array([-np.sin(ta), ecc + np.cos(ta), 0.0])
        cO, sO, ci, si, cw, sw = np.cos(raan), np.sin(raan), np.cos(inc), np.sin(inc), np.cos(aop), np.sin(aop)
        rot = np.array([[cO * cw - sO * sw * ci, -cO * sw - sO * cw * ci, sO * si],
                        [sO * cw + cO * sw * ci, -sO * sw + cO * cw * ci, -cO * si],
                        [sw * si, cw * si, ci]])
        rv, vv = rot @ rp, rot @ vp
        return cls(*rv.tolist(), *vv.tolist(), int(epoch_ns), frame)

    @property
    def radius_km(self) -> np.ndarray:
        return np.array([self.x_km, self.y_km, self.z_km])

    @property
    def velocity_km_s(self) -> np.ndarray:
        return np.array([self.vx_km_s, self.vy_km_s, self.vz_km_s])

    def to_cartesian_pos_vel(self) -> np.ndarray:
        return np.array([self.x_km, self.y_km, self.z_km, self.vx_km_s, self.vy_km_s, self.vz_km_s])

    def rmag_km(self) -> float:
        return float(np.sqrt((self.x_km * self.x_km + self.y_km * self.y_km) + self.z_km * self.z_km))

    @property
    def epoch(self) -> int:
        return self.epoch_ns


@dataclass(frozen=True)
class Mass:
    dry_mass_kg: float = 0.0
    prop_mass_kg: float = 0.0
    extra_mass_kg: float = 0.0

    def total_mass_kg(self) -> float:
        return self.dry_mass_kg + self.prop_mass_kg + self.extra_mass_kg


@dataclass(frozen=True)
class SRPData:
    area_m2: float = 0.0
    coeff_reflectivity: float = 1.8  # anise default


@dataclass(frozen=True)
class DragData:
    area_m2: float = 0.0
    coeff_drag: float = 2.2  # anise default


@dataclass(frozen=True)
class Spacecraft:
    """`Spacecraft` (cosmic/spacecraft.rs:115-143) without thruster / guidance / STM members."""

    orbit: Orbit
    mass: Mass = field(default_factory=Mass)
    srp: SRPData = field(default_factory=SRPData)
    drag: DragData = field(default_factory=DragData)

    @classmethod
    def from_orbit(cls, orbit: Orbit) -> "Spacecraft":
        return cls(orbit=orbit)

    def epoch(self) -> int:
        return self.orbit.epoch_ns

    def mass_kg(self) -> float:
        return self.mass.total_mass_kg()

    def to_vector(self) -> np.ndarray:
        """First 9 entries of `State::to_vector` (cosmic/spacecraft.rs:449-473)."""
        return np.array([*self.orbit.to_cartesian_pos_vel(), self.srp.coeff_reflectivity, self.drag.coeff_drag,
                         self.mass.prop_mass_kg])

    def with_vector(self, epoch_ns: int, vec: Sequence[float]) -> "Spacecraft":
        """`State::set` (cosmic/spacecraft.rs:477-497)."""
        o = Orbit(*[float(v) for v in vec[:6]], int(epoch_ns), self.orbit.frame)
        return replace(self, orbit=o, srp=replace(self.srp, coeff_reflectivity=float(vec[6])),
                       drag=replace(self.drag, coeff_drag=float(vec[7])),
                       mass=replace(self.mass, prop_mass_kg=float(vec[8])))


def pack_spacecraft(states: Iterable[Spacecraft]):
    """AoS `Vec<Spacecraft>` -> the SoA arrays of the C ABI: state[9][n], consts[4][n], epoch0[n]."""
    states = list(states)
    n = len(states)
    st = np.empty((9, n))
    cs = np.empty((4, n))
    ep = np.empty(n, dtype=np.int64)
    for i, sc in enumerate(states):
        st[:, i] = sc.to_vector()
        cs[:, i] = (sc.mass.dry_mass_kg, sc.mass.extra_mass_kg, sc.srp.area_m2, sc.drag.area_m2)
        ep[i] = sc.orbit.epoch_ns
    return st, cs, ep
