"""Config / wire formats of the hot path: build the engine's inputs from the reference's serde layouts.

* ``PropagatorConfig`` / ``Dynamics`` / ``AccelModels`` / ``ForceModels`` — dynamics/sequence/config.rs:96-169 (the closed
  set of models the GPU path accepts is exactly what these structs serialise);
* ``load_ground_stations`` — `GroundStation::load_named` on the YAML layout of examples/04_lro_od/dsn-network.yaml
  (od/ground_station/mod.rs:47-75: name, location{latitude_deg, longitude_deg, height_km, terrain_mask, frame},
  stochastic_noises{range_km, doppler_km_s}{white_noise{sigma}, bias{constant}}, measurement_types, light_time_correction);
* durations are hifitime strings ("1 min", "60 s", "2 h 30 min") or integer nanoseconds.

Dicts (already parsed YAML / JSON) and file paths are both accepted: `.dhall` files — the format the reference ships its
configurations in (data/02_config/prop_config.dhall, full_seq.dhall) — go through the data-subset reader `nyx_b200.dhall`,
anything else through PyYAML.  Nothing here touches the device.
"""
from __future__ import annotations

import re
from dataclasses import dataclass, field
from pathlib import Path
from typing import Dict, Optional, Union

from .dynamics import AtmDensity, Drag, DynamicsError, GravityField, OrbitalDynamics, PointMasses, SolarPressure, SpacecraftDynamics
from .frames import EARTH, IAU_EARTH_FRAME, IAU_MOON_FRAME, MOON, Almanac
from .gravity import GravityFieldData
from .od import GroundStation, MeasurementType, StochasticNoise
from .propagator import ErrorControl, IntegratorMethod, IntegratorOptions, Propagator

_UNITS_NS = {"ns": 1, "us": 10**3, "μs": 10**3, "ms": 10**6, "s": 10**9, "sec": 10**9, "min": 60 * 10**9, "h": 3600 * 10**9, "hr": 3600 * 10**9,
             "d": 86400 * 10**9, "day": 86400 * 10**9, "days": 86400 * 10**9}


def parse_duration(v) -> int:
    """hifitime `Duration` from its Display form ("1 min 30 s", "45 min", "0.5 s") or an integer number of nanoseconds."""
    if isinstance(v, (int,)):
        return int(v)
    if isinstance(v, float):
        return int(v * 1e9)  # bare floats are seconds
    total = 0
    toks = re.findall(r"([-+]?\d+(?:\.\d+)?(?:[eE][-+]?\d+)?)\s*([a-zA-Zμ]+)", str(v))
    if not toks:
        raise ValueError(f"cannot parse duration {v!r}")
    for num, unit in toks:
        if unit not in _UNITS_NS:
            raise ValueError(f"unknown duration unit {unit!r} in {v!r}")
        total += int(round(float(num) * _UNITS_NS[unit]))
    return total


def _load(src: Union[str, Path, dict]) -> dict:
    if isinstance(src, dict):
        return src
    if str(src).endswith(".dhall"):
        from . import dhall

        return dhall.load(src)
    import yaml

    return yaml.safe_load(Path(src).read_text())


_BODY_FIXED = {EARTH: IAU_EARTH_FRAME, MOON: IAU_MOON_FRAME}


def integrator_options_from(d: Optional[dict]) -> IntegratorOptions:
    """IntegratorOptions (propagators/options.rs:42-60); missing keys take the defaults of `IntegratorOptions::default`."""
    o = IntegratorOptions.default()
    if not d:
        return o
    for key in ("init_step", "min_step", "max_step"):
        if key in d:
            setattr(o, key, parse_duration(d[key]))
    if "tolerance" in d:
        o.tolerance = float(d["tolerance"])
    if "attempts" in d:
        o.attempts = int(d["attempts"])
    if "fixed_step" in d:
        o.fixed_step = bool(d["fixed_step"])
    if "error_ctrl" in d:
        o.error_ctrl = ErrorControl[d["error_ctrl"]] if isinstance(d["error_ctrl"], str) else ErrorControl(d["error_ctrl"])
    return o


@dataclass
class PropagatorConfig:
    """`PropagatorConfig{dynamics, method, options}` (dynamics/sequence/config.rs:137-152)."""

    dynamics: dict = field(default_factory=dict)
    method: IntegratorMethod = IntegratorMethod.RungeKutta89
    options: IntegratorOptions = field(default_factory=IntegratorOptions.default)

    @classmethod
    def load(cls, src: Union[str, Path, dict]) -> "PropagatorConfig":
        d = _load(src)
        m = d.get("method", "RungeKutta89")
        dyn = d.get("dynamics")
        if dyn is None:   # data/02_config/prop_config.dhall keeps the `Dynamics` fields beside `method` and `options`
            dyn = {k: d[k] for k in ("accel_models", "force_models") if k in d}
        return cls(dyn or {}, IntegratorMethod[m] if isinstance(m, str) else IntegratorMethod(m), integrator_options_from(d.get("options")))

    @classmethod
    def load_named(cls, src: Union[str, Path, dict]) -> Dict[str, "PropagatorConfig"]:
        """The `propagators` map of a sequence file (data/02_config/full_seq.dhall; serde_dhall writes maps as lists of
        `{ _1 = name, _2 = config }`), or a plain mapping name -> config."""
        from .dhall import pairs_to_dict

        d = _load(src)
        props = pairs_to_dict(d.get("propagators", d) if isinstance(d, dict) else d)
        return {str(name): cls.load(cfg) for name, cfg in props.items()}

    def build_dynamics(self, almanac: Optional[Almanac]) -> SpacecraftDynamics:
        """`Dynamics::build` (config.rs:104-134): two-body + [PointMasses] + [GravityField]; [SolarPressure], [Drag]."""
        am = self.dynamics.get("accel_models") or {}
        fm = self.dynamics.get("force_models") or {}
        accel = []
        if am.get("point_masses"):
            accel.append(PointMasses.new([int(b) for b in am["point_masses"]["celestial_objects"]]))
        if am.get("gravity_field"):
            g = am["gravity_field"]
            if "_1" in g:   # serde tuple (GravityFieldConfig, Frame uid) as the reference's Dhall files hold it
                g = dict(g["_1"], frame=g.get("_2") or {})
            fr = g.get("frame", {})
            frame = _BODY_FIXED.get(int(fr.get("ephemeris_id", EARTH)), IAU_EARTH_FRAME) if isinstance(fr, dict) else fr
            path = str(g["filepath"])
            # io/gravity.rs:99-115 routes on the extension (its `.cof` test never matches, App. B: we follow the intent)
            if ".cof" in path:
                gd = GravityFieldData.from_cof(path, int(g["degree"]), int(g["order"]), bool(g.get("gunzipped", path.endswith(".gz"))), frame)
            else:
                gd = GravityFieldData.from_shadr(path, int(g["degree"]), int(g["order"]), bool(g.get("gunzipped", path.endswith(".gz"))), frame)
            accel.append(GravityField.new(gd))
        if am.get("solid_tides"):
            raise DynamicsError("SolidTides is outside the closed set of models of the GPU path")
        orbital = OrbitalDynamics.new(accel) if accel else OrbitalDynamics.two_body()
        forces = []
        if fm.get("solar_pressure"):
            s = fm["solar_pressure"]
            if almanac is None:
                raise DynamicsError("planetary data from third body not loaded")
            shadows = [almanac.frames[int(b["ephemeris_id"]) if isinstance(b, dict) else int(b)]
                       for b in (s.get("shadow_model", {}) or {}).get("shadow_bodies", s.get("shadow_bodies", [EARTH]))]
            srp = SolarPressure.default_flux_raw(shadows, almanac)
            if "phi" in s:
                srp.phi = float(s["phi"])
            srp.estimate = bool(s.get("estimate", True))
            forces.append(srp)
        if fm.get("drag"):
            dg = fm["drag"]
            den = dg.get("density", "earth_exponential")
            if isinstance(den, dict):
                (kind, val), = den.items()
                density = {"Constant": lambda v: AtmDensity.Constant(float(v)),
                           "Exponential": lambda v: AtmDensity.Exponential(float(v["rho0"]), float(v["r0"]), float(v["ref_alt_m"])),
                           "StdAtm": lambda v: AtmDensity.StdAtm(float(v["max_alt_m"]))}[kind](val)
            else:
                density = AtmDensity.earth_exponential()
            forces.append(Drag(density, IAU_EARTH_FRAME, bool(dg.get("estimate", False))))
        if not forces:
            return SpacecraftDynamics.new(orbital)
        return SpacecraftDynamics.from_models(orbital, forces)

    def build(self, almanac: Optional[Almanac] = None, **kw) -> Propagator:
        """`PropagatorConfig::build` (config.rs:145-151)."""
        return Propagator.new(self.build_dynamics(almanac), self.method, self.options, **kw)


def _noise_from(d: Optional[dict]) -> StochasticNoise:
    if not d:
        raise ValueError("NoiseNotConfigured")
    wn = d.get("white_noise") or {}
    bias = (d.get("bias") or {})
    return StochasticNoise(float(wn.get("sigma", 0.0)), float(bias.get("constant", 0.0) or 0.0))


_MSR_KEYS = {"range_km": MeasurementType.Range, "doppler_km_s": MeasurementType.Doppler, "Range": MeasurementType.Range, "Doppler": MeasurementType.Doppler}


def load_ground_stations(src: Union[str, Path, dict]) -> Dict[str, GroundStation]:
    """`GroundStation::load_named`: a mapping name -> station in the layout of the reference's dsn-network.yaml."""
    out: Dict[str, GroundStation] = {}
    for key, g in _load(src).items():
        loc = g.get("location", g)
        fr = loc.get("frame", {}) or {}
        frame = _BODY_FIXED.get(int(fr.get("ephemeris_id", EARTH)))
        if frame is None:
            raise ValueError(f"{key}: no body-fixed frame model for ephemeris id {fr.get('ephemeris_id')}")
        mask = g.get("elevation_mask_deg")
        if mask is None:
            tm = loc.get("terrain_mask") or []
            mask = max((float(t.get("elevation_mask_deg", 0.0)) for t in tm), default=0.0)   # flat terrain: one entry
        if loc.get("terrain_mask_ignored"):
            mask = -90.0
        types = [_MSR_KEYS[t] for t in g.get("measurement_types", ["range_km", "doppler_km_s"]) if t in _MSR_KEYS]
        sn = g.get("stochastic_noises") or {}
        noises = {_MSR_KEYS[k]: _noise_from(v) for k, v in sn.items() if k in _MSR_KEYS}
        it = g.get("integration_time")
        out[key] = GroundStation(name=g.get("name", key), latitude_deg=float(loc["latitude_deg"]), longitude_deg=float(loc["longitude_deg"]),
                                 height_km=float(loc["height_km"]), frame=frame, elevation_mask_deg=float(mask), measurement_types=types,
                                 stochastic_noises=noises, integration_time=parse_duration(it) if it else None,
                                 light_time_correction=bool(g.get("light_time_correction", False)))
    return out
