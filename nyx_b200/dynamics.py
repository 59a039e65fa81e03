"""Force-model descriptors: host-side mirror of ``dynamics/*.rs`` for the propagation path.

The reference's `SpacecraftDynamics` is an *open* set of `Arc<dyn AccelModel>` /
`Arc<dyn ForceModel>` trait objects (spacecraft.rs:44-49, orbital.rs:44-46).  The GPU
engine accepts the *closed* set that `dynamics/sequence/config.rs:96-169` serialises:
two-body (always on, orbital.rs:86-92) + PointMasses + GravityField + SolarPressure + Drag.
``SpacecraftDynamics.pack()`` lowers them to the C-ABI PODs of ``include/nyxb.h``.
No arithmetic of the hot path lives here: that is all in ``csrc/`` (CUDA).
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field
from typing import List, Optional, Sequence

import numpy as np

from . import abi
from .frames import Almanac, Frame, Rotation, SUN_J2000
from .gravity import GravityFieldData

SOLAR_FLUX_W_m2 = 1367.0  # solarpressure.rs:35


class DynamicsError(RuntimeError):
    """`DynamicsError` (dynamics/mod.rs:177-203)."""


# --------------------------------------------------------------------------- accel models
@dataclass
class PointMasses:
    """`PointMasses` (orbital.rs:174-247): third-body point-mass gravity."""

    celestial_objects: List[int]

    @classmethod
    def new(cls, celestial_objects: Sequence[int]) -> "PointMasses":
        return cls(list(celestial_objects))


@dataclass
class GravityField:
    """`GravityField` (gravity_field.rs:36-132): normalised spherical harmonics."""

    grav_data: GravityFieldData

    @classmethod
    def new(cls, stor: GravityFieldData) -> "GravityField":
        return cls(stor)


@dataclass
class OrbitalDynamics:
    """`OrbitalDynamics` (orbital.rs:44-78). Two-body gravity is always included."""

    accel_models: list = field(default_factory=list)

    @classmethod
    def two_body(cls) -> "OrbitalDynamics":
        return cls([])

    @classmethod
    def point_masses(cls, celestial_objects: Sequence[int]) -> "OrbitalDynamics":
        return cls([PointMasses.new(celestial_objects)])

    @classmethod
    def from_model(cls, accel_model) -> "OrbitalDynamics":
        return cls([accel_model])

    @classmethod
    def new(cls, accel_models: Sequence) -> "OrbitalDynamics":
        return cls(list(accel_models))


# --------------------------------------------------------------------------- force models
@dataclass
class ShadowModel:
    """`ShadowModel` (cosmic/eclipse.rs:35-83)."""

    light_source: Frame
    shadow_bodies: List[Frame]


@dataclass
class SolarPressure:
    """`SolarPressure` (solarpressure.rs:43-165)."""

    phi: float
    shadow_model: ShadowModel
    estimate: bool = False

    @classmethod
    def default_flux_raw(cls, shadow_bodies: Sequence[Frame], almanac: Almanac) -> "SolarPressure":
        return cls(SOLAR_FLUX_W_m2, ShadowModel(almanac.frame_info(SUN_J2000),
                                                 [almanac.frame_info(b) for b in shadow_bodies]), True)

    @classmethod
    def new(cls, shadow_bodies: Sequence[Frame], almanac: Almanac) -> "SolarPressure":
        return cls.default_flux_raw(shadow_bodies, almanac)

    @classmethod
    def default_flux(cls, shadow_body: Frame, almanac: Almanac) -> "SolarPressure":
        return cls.default_flux_raw([shadow_body], almanac)

    @classmethod
    def default_no_estimation(cls, shadow_bodies: Sequence[Frame], almanac: Almanac) -> "SolarPressure":
        srp = cls.default_flux_raw(shadow_bodies, almanac)
        srp.estimate = False
        return srp

    @classmethod
    def with_flux(cls, flux_w_m2: float, shadow_bodies: Sequence[Frame], almanac: Almanac) -> "SolarPressure":
        srp = cls.default_flux_raw(shadow_bodies, almanac)
        srp.phi = flux_w_m2
        return srp


@dataclass(frozen=True)
class AtmDensity:
    """`AtmDensity` (drag.rs:36-42): Constant(rho) | Exponential{rho0,r0,ref_alt_m} | StdAtm{max_alt_m}."""

    kind: int
    rho0: float = 0.0
    r0: float = 0.0
    ref_alt_m: float = 0.0

    @classmethod
    def Constant(cls, rho: float) -> "AtmDensity":
        return cls(abi.DENSITY_CONSTANT, rho0=rho)

    @classmethod
    def Exponential(cls, rho0: float, r0: float, ref_alt_m: float) -> "AtmDensity":
        return cls(abi.DENSITY_EXPONENTIAL, rho0=rho0, r0=r0, ref_alt_m=ref_alt_m)

    @classmethod
    def StdAtm(cls, max_alt_m: float) -> "AtmDensity":
        return cls(abi.DENSITY_STDATM, ref_alt_m=max_alt_m)

    @classmethod
    def earth_exponential(cls) -> "AtmDensity":
        return cls.Exponential(3.614e-13, 700_000.0, 88_667.0)  # drag.rs:134-148


@dataclass
class Drag:
    """`Drag` (drag.rs:123-284); `ConstantDrag` (drag.rs:68-107) == Drag with AtmDensity.Constant."""

    density: AtmDensity
    frame: Frame
    estimate: bool = False

    @classmethod
    def earth_exp(cls, almanac: Almanac) -> "Drag":
        from .frames import IAU_EARTH_FRAME

        return cls(AtmDensity.earth_exponential(), almanac.frame_info(IAU_EARTH_FRAME))

    @classmethod
    def std_atm1976(cls, almanac: Almanac) -> "Drag":
        from .frames import IAU_EARTH_FRAME

        return cls(AtmDensity.StdAtm(1_000_000.0), almanac.frame_info(IAU_EARTH_FRAME))


# --------------------------------------------------------------------------- spacecraft dynamics
def _rotation_c(rot: Optional[Rotation]) -> abi.Rotation:
    if rot is None or rot.kind == 0:
        return abi.Rotation(kind=0)
    return abi.Rotation(1, 0, rot.ra0_deg, rot.ra1_deg_cy, rot.dec0_deg, rot.dec1_deg_cy, rot.w0_deg, rot.w1_deg_day)


@dataclass
class PackedDynamics:
    """A `nyxb_dynamics` POD plus every buffer it points to (kept alive together)."""

    c: abi.DynamicsC
    keep: list

    def byref(self):
        return C.byref(self.c)


@dataclass
class SpacecraftDynamics:
    """`SpacecraftDynamics` (spacecraft.rs:44-135) restricted to coast arcs (guid_law = None)."""

    orbital_dyn: OrbitalDynamics
    force_models: list = field(default_factory=list)

    @classmethod
    def new(cls, orbital_dyn: OrbitalDynamics) -> "SpacecraftDynamics":
        return cls(orbital_dyn, [])

    @classmethod
    def from_model(cls, orbital_dyn: OrbitalDynamics, force_model) -> "SpacecraftDynamics":
        return cls(orbital_dyn, [force_model])

    @classmethod
    def from_models(cls, orbital_dyn: OrbitalDynamics, force_models: Sequence) -> "SpacecraftDynamics":
        return cls(orbital_dyn, list(force_models))

    # ------------------------------------------------------------------ lowering to the C ABI
    def pack(self, frame: Frame, almanac: Optional[Almanac]) -> PackedDynamics:
        """Lower to `nyxb_dynamics` for states expressed in `frame` (the integration frame)."""
        keep: list = []
        dyn = abi.DynamicsC()
        dyn.mu_central_km3_s2 = frame.mu_km3_s2()  # orbital.rs:86-90: the *state's* frame mu
        dyn.central_radius_km = frame.radius_km if frame.radius_km is not None else 0.0

        pm: Optional[PointMasses] = None
        fields: List[GravityField] = []
        for model in self.orbital_dyn.accel_models:
            if isinstance(model, PointMasses):
                if pm is not None:
                    raise DynamicsError("only one PointMasses model is supported on the GPU path")
                pm = model
            elif isinstance(model, GravityField):
                if len(fields) >= abi.NYXB_MAX_FIELDS:
                    raise DynamicsError(f"at most {abi.NYXB_MAX_FIELDS} GravityField models are supported on the GPU path")
                fields.append(model)
            else:
                raise DynamicsError(f"unsupported acceleration model {type(model).__name__} (closed set only)")
        srp: Optional[SolarPressure] = None
        drag: Optional[Drag] = None
        for model in self.force_models:
            if isinstance(model, SolarPressure):
                srp = model
            elif isinstance(model, Drag):
                drag = model
            else:
                raise DynamicsError(f"unsupported force model {type(model).__name__} (closed set only)")

        # ---- bodies: every ephemeris the almanac holds is made available.  The tables are positions RELATIVE TO THE ALMANAC'S
        # CENTRE; the kernels use them as positions relative to the integration-frame centre (the reference re-centres through
        # anise's `almanac.transform(third_body_frame, osc.frame, ...)`, orbital.rs:230-234), so the two must be the same body.
        if almanac is not None and almanac.bodies:
            if almanac.center.ephemeris_id != frame.ephemeris_id:
                raise DynamicsError(f"almanac is centred on {almanac.center.name} but the states are expressed in {frame.name}: "
                                    f"build it with Almanac.synthetic(center=<the integration frame>)")
            if any(b.frame.ephemeris_id == frame.ephemeris_id for b in almanac.bodies):
                raise DynamicsError(f"almanac holds an ephemeris of its own centre ({frame.name})")
        bodies = almanac.bodies if almanac is not None else []
        if len(bodies) > abi.NYXB_MAX_BODIES:
            raise DynamicsError("too many ephemeris bodies")
        if bodies:
            arr = (abi.BodyC * len(bodies))()
            for i, b in enumerate(bodies):
                co = np.ascontiguousarray(b.coeffs, dtype=np.float64)
                keep.append(co)
                arr[i] = abi.BodyC(b.frame.mu_km3_s2(), b.frame.radius_km or 0.0, b.t0_ns, b.interval_ns,
                                   b.n_intervals, b.n_coeffs, abi.as_double_p(co))
            keep.append(arr)
            dyn.n_bodies = len(bodies)
            dyn.bodies = C.cast(arr, C.POINTER(abi.BodyC))

        if pm is not None:
            mask, k = 0, 0
            for obj in pm.celestial_objects:   # summation order of PointMasses::eom (orbital.rs:217)
                if obj == frame.ephemeris_id:
                    continue  # orbital.rs:219-222: the central body is handled by the two-body term
                if almanac is None:
                    raise DynamicsError("planetary data from third body not loaded")
                j = almanac.body_index(obj)
                if not (mask >> j) & 1:
                    dyn.point_mass_order[k] = j
                    k += 1
                mask |= 1 << j
            dyn.point_mass_mask = mask
            dyn.n_point_masses = k

        if fields:
            # every harmonic field is evaluated in ITS OWN body's frame (gravity_field.rs:149-154): a field of another body than the
            # integration centre needs that body's ephemeris.  The first field is the one the cooperative kernels parallelise.
            arr = (abi.GravityFieldC * len(fields))()
            for i, gf in enumerate(fields):
                gd = gf.grav_data
                n = gd.degree
                c = np.ascontiguousarray(gd.c_nm[: n + 1, : n + 1], dtype=np.float64)
                s = np.ascontiguousarray(gd.s_nm[: n + 1, : n + 1], dtype=np.float64)
                if gd.frame.ephemeris_id == frame.ephemeris_id:
                    body = abi.NYXB_CENTRAL_BODY
                else:
                    if almanac is None or not almanac.has_body(gd.frame.ephemeris_id):
                        raise DynamicsError(f"planetary data of the {gd.frame.name} gravity field's body not loaded")
                    body = almanac.body_index(gd.frame.ephemeris_id)
                arr[i] = abi.GravityFieldC(n, gd.order, gd.frame.mu_km3_s2(), gd.frame.mean_equatorial_radius_km(),
                                           abi.as_double_p(c), abi.as_double_p(s), _rotation_c(gd.frame.rotation), body, 0)
                keep += [c, s]
            keep.append(arr)
            dyn.n_gravity = len(fields)
            dyn.gravity = C.cast(arr, C.POINTER(abi.GravityFieldC))

        if srp is not None:
            if almanac is None:
                raise DynamicsError("planetary data from third body not loaded")
            sc = abi.SrpC()
            sc.phi_w_m2 = srp.phi
            sc.sun_body = almanac.body_index(srp.shadow_model.light_source.ephemeris_id)
            if len(srp.shadow_model.shadow_bodies) > 4:
                raise DynamicsError("at most 4 shadow bodies")
            sc.n_shadow = len(srp.shadow_model.shadow_bodies)
            sc.estimate = 1 if srp.estimate else 0
            for q, fb in enumerate(srp.shadow_model.shadow_bodies):
                sc.shadow_body[q] = (abi.NYXB_CENTRAL_BODY if fb.ephemeris_id == frame.ephemeris_id
                                     else almanac.body_index(fb.ephemeris_id))
            keep.append(sc)
            dyn.srp = C.pointer(sc)

        if drag is not None:
            dc = abi.DragC(drag.density.kind, 0, drag.density.rho0, drag.density.r0, drag.density.ref_alt_m,
                           drag.frame.mean_equatorial_radius_km(), _rotation_c(drag.frame.rotation))
            keep.append(dc)
            dyn.drag = C.pointer(dc)

        return PackedDynamics(dyn, keep)
