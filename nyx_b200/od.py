"""Host-side mirror of the reference's sequential-filter orbit determination surface for the batched GPU path
(BASELINE configs[4]).  Reference paths are relative to /root/reference/nyx-core/src:

* ``GroundStation``            od/ground_station/{mod.rs:47-75, builtin.rs:25-117, trk_device.rs:35-253}
* ``MeasurementType``          od/msr/types.rs:31-45
* ``TrackingDataArc``          od/msr/trackingdata (epochs + tracker + data per type), reduced to arrays
* ``ProcessNoise3D``           od/snc.rs:38-56, 118-134, 288-311
* ``SigmaRejection``           od/process/rejectcrit.rs:35-46
* ``KfEstimate``               od/estimate/kfestimate.rs (nominal state, covariance, state deviation)
* ``SpacecraftUncertainty``    od/estimate/sc_uncertainty.rs:36-138
* ``KalmanODProcess``          od/process/{initializers.rs:60-113, mod.rs:128-497}; `SpacecraftKalmanOD` = MsrSize 2,
                               `SpacecraftKalmanScalarOD` = MsrSize 1 (od/mod.rs:77-91)

Nothing here runs the filter: ``KalmanODProcess.process_arcs`` packs the ensemble into the SoA arrays of
``nyxb_od_ekf_batch`` (include/nyxb.h) — ONE kernel launch runs every filter from the first to the last measurement
(propagation with the STM, time updates, measurement updates, state replacement) on the device.  There is no CPU
fallback.  The measurement *simulator* (`simulate_tracking`, stands in for od/simulator/arc.rs) is host-side data
generation for tests and the benchmark, not part of the filter path.
"""
from __future__ import annotations

import enum
import math
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence

import numpy as np

from . import abi
from .cosmic import Orbit, Spacecraft, duration_to_seconds
from .dynamics import _rotation_c
from .frames import IAU_EARTH_FRAME, NS_PER_S, Almanac, Frame


class MeasurementType(enum.IntEnum):
    Range = abi.MSR_RANGE
    Doppler = abi.MSR_DOPPLER


class KalmanVariant(enum.IntEnum):
    ReferenceUpdate = abi.KF_REFERENCE_UPDATE      # EKF
    DeviationTracking = abi.KF_DEVIATION_TRACKING  # CKF


class LocalFrame(enum.IntEnum):
    Inertial = 0
    RIC = 1


class ODError(RuntimeError):
    pass


@dataclass(frozen=True)
class StochasticNoise:
    """`StochasticNoise` reduced to what the filter reads: the white-noise sigma (covariance = sigma^2,
    noise/white.rs) and the constant part of the bias (trk_device.rs:238-253)."""

    sigma: float
    bias_constant: float = 0.0

    def covariance(self, _epoch_ns: int = 0) -> float:
        return self.sigma ** 2

    @classmethod
    def default_range_km(cls) -> "StochasticNoise":
        return cls(2e-3)     # noise/mod.rs: 2 m

    @classmethod
    def default_doppler_km_s(cls) -> "StochasticNoise":
        return cls(3e-6)     # noise/mod.rs: 3 mm/s


@dataclass
class GroundStation:
    name: str
    latitude_deg: float
    longitude_deg: float
    height_km: float
    frame: Frame = IAU_EARTH_FRAME
    elevation_mask_deg: float = 0.0
    measurement_types: Sequence[MeasurementType] = (MeasurementType.Range, MeasurementType.Doppler)
    stochastic_noises: Dict[MeasurementType, StochasticNoise] = field(default_factory=lambda: {
        MeasurementType.Range: StochasticNoise.default_range_km(), MeasurementType.Doppler: StochasticNoise.default_doppler_km_s()})
    integration_time: Optional[int] = None
    light_time_correction: bool = False

    # builtin.rs:25-117
    @classmethod
    def dss65_madrid(cls, elevation_mask_deg, range_noise_km: StochasticNoise, doppler_noise_km_s: StochasticNoise):
        return cls("Madrid", 40.427_222, 4.250_556, 0.834_939, IAU_EARTH_FRAME, elevation_mask_deg,
                   stochastic_noises={MeasurementType.Range: range_noise_km, MeasurementType.Doppler: doppler_noise_km_s})

    @classmethod
    def dss34_canberra(cls, elevation_mask_deg, range_noise_km: StochasticNoise, doppler_noise_km_s: StochasticNoise):
        return cls("Canberra", -35.398_333, 148.981_944, 0.691_750, IAU_EARTH_FRAME, elevation_mask_deg,
                   stochastic_noises={MeasurementType.Range: range_noise_km, MeasurementType.Doppler: doppler_noise_km_s})

    @classmethod
    def dss13_goldstone(cls, elevation_mask_deg, range_noise_km: StochasticNoise, doppler_noise_km_s: StochasticNoise):
        return cls("Goldstone", 35.247_164, 243.205, 1.071_149_04, IAU_EARTH_FRAME, elevation_mask_deg,
                   stochastic_noises={MeasurementType.Range: range_noise_km, MeasurementType.Doppler: doppler_noise_km_s})

    def body_fixed(self):
        """Geodetic (lat, long, height) -> body-fixed Cartesian position and local zenith on the frame's ellipsoid
        (anise `Orbit::try_latlongalt`; sphere when the frame has no polar radius)."""
        a = self.frame.mean_equatorial_radius_km()
        b = self.frame.polar_radius_km if self.frame.polar_radius_km is not None else a
        e2 = 1.0 - (b * b) / (a * a)
        lat, lon = math.radians(self.latitude_deg), math.radians(self.longitude_deg)
        sl, cl = math.sin(lat), math.cos(lat)
        nu = a / math.sqrt(1.0 - e2 * sl * sl)
        pos = np.array([(nu + self.height_km) * cl * math.cos(lon), (nu + self.height_km) * cl * math.sin(lon),
                        (nu * (1.0 - e2) + self.height_km) * sl])
        up = np.array([cl * math.cos(lon), cl * math.sin(lon), sl])
        return pos, up

    def to_c(self, integration_frame: Frame, almanac: Optional[Almanac]) -> abi.GroundStationC:
        if self.integration_time is not None or self.light_time_correction:
            raise ODError("only instantaneous measurements without light-time correction are supported on the GPU path")
        types = list(self.measurement_types)
        if not 1 <= len(types) <= 2 or len(set(types)) != len(types):
            raise ODError("a ground station carries one or two of {Range, Doppler}")
        g = abi.GroundStationC()
        pos, up = self.body_fixed()
        for i in range(3):
            g.pos_fixed_km[i] = pos[i]
            g.up_fixed[i] = up[i]
        g.elevation_mask_deg = self.elevation_mask_deg
        g.rot = _rotation_c(self.frame.rotation)
        if self.frame.ephemeris_id == integration_frame.ephemeris_id:
            g.body = abi.NYXB_CENTRAL_BODY
            g.body_radius_km = -1.0   # same body: the elevation mask is the only visibility test (trk_device.rs:162-166)
        else:
            if almanac is None:
                raise ODError("an almanac with the station's body is needed when it does not sit on the integration centre")
            g.body = almanac.body_index(self.frame.ephemeris_id)
            g.body_radius_km = integration_frame.mean_equatorial_radius_km()
        g.n_types = len(types)
        for i, t in enumerate(types):
            if t not in self.stochastic_noises:
                raise ODError(f"NoiseNotConfigured: {t.name}")
            g.types[i] = int(t)
            g.noise_var[i] = self.stochastic_noises[t].covariance()
            g.bias[i] = self.stochastic_noises[t].bias_constant
        return g


@dataclass
class TrackingDataArc:
    """One tracking schedule (epochs + tracker names) with `n` observation sets: obs[k][type][i], NaN = type not in
    the measurement's data (both NaN: measurement k absent from arc i)."""

    epoch_ns: np.ndarray            # [m] int64 ascending
    tracker: List[str]              # [m]
    obs: np.ndarray                 # [m][2][n] float64

    def __post_init__(self):
        self.epoch_ns = np.ascontiguousarray(self.epoch_ns, dtype=np.int64)
        self.obs = np.ascontiguousarray(self.obs, dtype=np.float64)
        m = self.epoch_ns.shape[0]
        if self.obs.ndim == 2:
            self.obs = np.ascontiguousarray(self.obs[:, :, None])
        if len(self.tracker) != m or self.obs.shape[0] != m or self.obs.shape[1] != 2:
            raise ODError("expected epoch_ns[m], tracker[m], obs[m][2][n]")
        if m and np.any(np.diff(self.epoch_ns) < 0):
            raise ODError("measurement epochs must be ascending")

    def __len__(self):
        return self.epoch_ns.shape[0]

    @property
    def n(self) -> int:
        return self.obs.shape[2]

    # ---- parquet I/O in the reference's layout (od/msr/trackingdata/io_parquet.rs:43-354): one arc per file
    _COLUMNS = ("Range (km)", "Doppler (km/s)")   # MeasurementType::to_field names (od/msr/types.rs)

    def to_parquet(self, path, index: int = 0, metadata: Optional[dict] = None):
        """`TrackingDataArc::to_parquet` for the observation set `index`: "Epoch (UTC)", "Tracking device" and one nullable
        Float64 column per measurement type present; measurements absent from this arc are not written."""
        import pyarrow as pa
        import pyarrow.parquet as pq

        from .cosmic import epochs_to_utc_iso

        o = self.obs[:, :, index]
        present = ~np.isnan(o).all(axis=1)
        if not present.any():
            raise ODError("EmptyDataset: tracking data arc to parquet")
        cols = [pa.array(epochs_to_utc_iso(self.epoch_ns[present]), type=pa.string()),
                pa.array([t for t, p in zip(self.tracker, present) if p], type=pa.string())]
        fields = [pa.field("Epoch (UTC)", pa.string(), nullable=False), pa.field("Tracking device", pa.string(), nullable=False)]
        for c, name in enumerate(self._COLUMNS):
            v = o[present, c]
            if np.isnan(v).all():
                continue   # unique_types(): a type no measurement carries has no column
            cols.append(pa.array(v, type=pa.float64(), mask=np.isnan(v)))
            fields.append(pa.field(name, pa.float64(), nullable=True, metadata={"unit": name[name.index("(") + 1:-1]}))
        meta = {"Purpose": "Tracking Arc Data"}
        meta.update(metadata or {})
        pq.write_table(pa.Table.from_arrays(cols, schema=pa.schema(fields, metadata=meta)), str(path))
        return path

    @classmethod
    def from_parquet(cls, path) -> "TrackingDataArc":
        """`TrackingDataArc::from_parquet` (io_parquet.rs:43-213): needs "Epoch (UTC)", "Tracking device" and at least one of
        the measurement columns this path knows (range, Doppler); rows are sorted by epoch."""
        import pyarrow.parquet as pq

        from .cosmic import utc_iso_to_epochs

        tab = pq.read_table(str(path))
        names = set(tab.column_names)
        for need in ("Epoch (UTC)", "Tracking device"):
            if need not in names:
                raise ODError(f"MissingData: {need}")
        if not names & set(cls._COLUMNS):
            raise ODError("MissingData: `Range (km)` or `Doppler (km/s)`")
        ep = utc_iso_to_epochs(tab["Epoch (UTC)"].to_pylist())
        obs = np.full((len(ep), 2), np.nan)
        for c, name in enumerate(cls._COLUMNS):
            if name in names:
                obs[:, c] = [np.nan if v is None else v for v in tab[name].to_pylist()]
        order = np.argsort(ep, kind="stable")
        trk = tab["Tracking device"].to_pylist()
        return cls(ep[order], [trk[i] for i in order], obs[order][:, :, None])

    @classmethod
    def stack(cls, arcs: Sequence["TrackingDataArc"]) -> "TrackingDataArc":
        """n single-observation-set arcs -> one arc with n observation sets over the union of their schedules (what
        `process_arcs` takes); a (epoch, tracker) pair missing from an arc is NaN there."""
        keys = sorted({(int(e), t) for a in arcs for e, t in zip(a.epoch_ns, a.tracker)})
        pos = {k: i for i, k in enumerate(keys)}
        n = sum(a.n for a in arcs)
        obs = np.full((len(keys), 2, n), np.nan)
        col = 0
        for a in arcs:
            rows = [pos[(int(e), t)] for e, t in zip(a.epoch_ns, a.tracker)]
            obs[rows, :, col:col + a.n] = a.obs
            col += a.n
        return cls(np.array([k[0] for k in keys], dtype=np.int64), [k[1] for k in keys], obs)


@dataclass(frozen=True)
class SigmaRejection:
    num_sigmas: float = 3.0


@dataclass
class ProcessNoise3D:
    diag: np.ndarray
    disable_time: int
    local_frame: Optional[LocalFrame] = None

    @classmethod
    def from_diagonal(cls, values, disable_time: int, local_frame: Optional[LocalFrame] = None):
        v = np.asarray(values, dtype=np.float64)
        assert v.shape == (3,), "Not enough values for the size of the SNC matrix"
        return cls(v, int(disable_time), local_frame)

    @classmethod
    def from_velocity_km_s(cls, velocity_noise, noise_duration: int, disable_time: int, local_frame: Optional[LocalFrame] = None):
        """snc.rs:288-311: diag = velocity noise / noise duration (seconds)."""
        return cls(np.asarray(velocity_noise, dtype=np.float64) / duration_to_seconds(noise_duration), int(disable_time), local_frame)


def dcm_ric_to_inertial(orbit: Orbit) -> np.ndarray:
    """Columns R, I, C (anise `Orbit::dcm_to_inertial(LocalFrame::RIC)`): R = r/|r|, C = h/|h|, I = C x R."""
    r, v = orbit.radius_km, orbit.velocity_km_s
    rh = r / np.linalg.norm(r)
    h = np.cross(r, v)
    ch = h / np.linalg.norm(h)
    ih = np.cross(ch, rh)
    return np.column_stack([rh, ih, ch])


@dataclass
class KfEstimate:
    nominal_state: Spacecraft
    covar: np.ndarray                       # [9][9]
    state_deviation: np.ndarray = field(default_factory=lambda: np.zeros(9))

    @classmethod
    def from_covar(cls, nominal_state: Spacecraft, covar) -> "KfEstimate":
        return cls(nominal_state, np.array(covar, dtype=np.float64).reshape(9, 9))

    @classmethod
    def from_diag(cls, nominal_state: Spacecraft, diag) -> "KfEstimate":
        return cls(nominal_state, np.diag(np.asarray(diag, dtype=np.float64)))

    def state(self) -> Spacecraft:
        """nominal + deviation (`Spacecraft + OVector<9>`, cosmic/spacecraft.rs:713-728: Cr clamped to [0, 2])."""
        v = self.nominal_state.to_vector() + self.state_deviation
        v[6] = min(max(v[6], 0.0), 2.0)
        return self.nominal_state.with_vector(self.nominal_state.epoch(), v)


@dataclass
class SpacecraftUncertainty:
    """sc_uncertainty.rs:36-138 (defaults included).  As coded the covariance is rotated as D^T C D with D = the
    local->inertial state DCM; the rotation-rate block of the RIC state DCM is taken as zero here (anise's
    `rot_mat_dt` is not in the tree)."""

    nominal: Spacecraft
    frame: Optional[LocalFrame] = None
    x_km: float = 0.5
    y_km: float = 0.5
    z_km: float = 0.5
    vx_km_s: float = 50e-5
    vy_km_s: float = 50e-5
    vz_km_s: float = 50e-5
    coeff_reflectivity: float = 0.0
    coeff_drag: float = 0.0
    mass_kg: float = 0.0

    def to_estimate(self) -> KfEstimate:
        vals = [self.x_km, self.y_km, self.z_km, self.vx_km_s, self.vy_km_s, self.vz_km_s, self.coeff_reflectivity,
                self.coeff_drag, self.mass_kg]
        if any(v < 0.0 for v in vals):
            raise ODError("uncertainties must be positive")
        d3 = dcm_ric_to_inertial(self.nominal.orbit) if self.frame == LocalFrame.RIC else np.eye(3)
        d6 = np.zeros((6, 6))
        d6[:3, :3] = d3
        d6[3:, 3:] = d3
        cov = np.zeros((9, 9))
        cov[:6, :6] = d6.T @ np.diag(np.square(vals[:6])) @ d6
        for i in range(6, 9):
            cov[i, i] = vals[i] ** 2
        return KfEstimate.from_covar(self.nominal, cov)


@dataclass
class ODSolution:
    """Results of n filters: final estimates and the per-measurement residual records (od/process/solution)."""

    final_state_soa: np.ndarray      # [9][n]
    final_epoch_ns: np.ndarray       # [n]
    covar: np.ndarray                # [n][9][9]
    state_deviation: np.ndarray      # [9][n]
    resid_ratio: np.ndarray          # [m][2][n]
    prefit: np.ndarray               # [m][2][n]
    postfit: np.ndarray              # [m][2][n]
    msr_flags: np.ndarray            # [m][n]
    est_state: Optional[np.ndarray]  # [m][9][n]
    est_covar_diag: Optional[np.ndarray]
    details: np.ndarray
    status: np.ndarray
    templates: Sequence[Spacecraft] = ()
    arc: Optional["TrackingDataArc"] = None
    devices: Optional[Dict[str, GroundStation]] = None

    def accepted(self) -> np.ndarray:
        return ((self.msr_flags & abi.MSRF_PROCESSED) != 0) & ((self.msr_flags & abi.MSRF_REJECTED) == 0)

    def rejected(self) -> np.ndarray:
        return (self.msr_flags & abi.MSRF_REJECTED) != 0

    def final_estimate(self, i: int) -> KfEstimate:
        sc = self.templates[i].with_vector(int(self.final_epoch_ns[i]), self.final_state_soa[:, i])
        return KfEstimate(sc, self.covar[i].copy(), self.state_deviation[:, i].copy())

    def to_parquet(self, path, index: int = 0, fields=None, metadata: Optional[dict] = None):
        """`ODSolution::to_parquet` (od/process/solution/export.rs:60-688) for filter `index`, one row per processed measurement
        epoch, with the columns this path records: "Epoch (UTC)", the state parameters of the estimate (default
        `Spacecraft::export_params`), "Sigma <item> (<frame>) (<unit>)" from the covariance diagonal, prefit / postfit residuals
        per measurement type, "Residual ratio", "Residual Rejected", "Tracker".  The full covariance, RIC sigmas, gains and
        filter-smoother ratios of the reference's file need the off-diagonal terms per epoch, which the kernel does not write
        back.  Needs `process_arcs(.., record_estimates=True)`."""
        import pyarrow as pa
        import pyarrow.parquet as pq

        from .cosmic import epochs_to_utc_iso
        from .param import EXPORT_PARAMS, StateError, evaluate

        if self.est_state is None or self.est_covar_diag is None or self.arc is None:
            raise ODError("no per-measurement estimates recorded: run process_arcs(.., record_estimates=True)")
        rows = np.nonzero((self.msr_flags[:, index] & abi.MSRF_PROCESSED) != 0)[0]
        if len(rows) == 0:
            raise ODError("EmptyDataset: no measurement was processed")
        tmpl = self.templates[index]
        frame = tmpl.orbit.frame
        est = self.est_state[rows, :, index].T          # [9][k]
        cols = [pa.array(epochs_to_utc_iso(self.arc.epoch_ns[rows]), type=pa.string())]
        schema = [pa.field("Epoch (UTC)", pa.string(), nullable=False)]
        for f in (EXPORT_PARAMS if fields is None else fields):
            try:
                vals = evaluate(f, est[:6], frame.mu_km3_s2(), tmpl, cr=est[6], cd=est[7], prop_mass_kg=est[8])
            except StateError:
                continue
            cols.append(pa.array(vals, type=pa.float64()))
            schema.append(pa.field(str(f), pa.float64(), nullable=False, metadata={"unit": f.unit, "Frame": frame.name}))
        items = ("X", "Y", "Z", "Vx", "Vy", "Vz", "Cr", "Cd", "Mass")
        units = ("km", "km", "km", "km/s", "km/s", "km/s", "unitless", "unitless", "kg")
        sig = np.sqrt(np.maximum(self.est_covar_diag[rows, :, index], 0.0))
        for q, (it, un) in enumerate(zip(items, units)):
            cols.append(pa.array(sig[:, q], type=pa.float64()))
            schema.append(pa.field(f"Sigma {it} ({frame.name}) ({un})", pa.float64(), nullable=False))
        # residual slots follow the order of the tracker's measurement types (include/nyxb.h: nyxb_od_outputs)
        slot_type = np.full((len(rows), 2), -1)
        for a, r in enumerate(rows):
            dev = (self.devices or {}).get(self.arc.tracker[r])
            types = list(dev.measurement_types) if dev is not None else [MeasurementType.Range, MeasurementType.Doppler]
            slot_type[a, :len(types)] = [int(t) for t in types]
        for label, arr in (("Prefit residual", self.prefit), ("Postfit residual", self.postfit)):
            for mt, un in ((MeasurementType.Range, "km"), (MeasurementType.Doppler, "km/s")):
                v = np.full(len(rows), np.nan)
                for q in range(2):
                    hit = slot_type[:, q] == int(mt)
                    v[hit] = arr[rows[hit], q, index]
                cols.append(pa.array(v, type=pa.float64(), mask=np.isnan(v)))
                schema.append(pa.field(f"{label}: {mt.name} ({un})", pa.float64(), nullable=True))
        ratio = self.resid_ratio[rows, 0, index]
        ratio = np.where(np.isnan(ratio), self.resid_ratio[rows, 1, index], ratio)
        cols.append(pa.array(ratio, type=pa.float64(), mask=np.isnan(ratio)))
        schema.append(pa.field("Residual ratio", pa.float64(), nullable=True))
        cols.append(pa.array((self.msr_flags[rows, index] & abi.MSRF_REJECTED) != 0, type=pa.bool_()))
        schema.append(pa.field("Residual Rejected", pa.bool_(), nullable=True))
        cols.append(pa.array([self.arc.tracker[r] for r in rows], type=pa.string()))
        schema.append(pa.field("Tracker", pa.string(), nullable=True))
        meta = {"Purpose": "Orbit determination results"}
        meta.update(metadata or {})
        pq.write_table(pa.Table.from_arrays(cols, schema=pa.schema(schema, metadata=meta)), str(path))
        return path


class KalmanODProcess:
    """`KalmanODProcess<SpacecraftDynamics, MsrSize, 3, GroundStation>` (od/process/mod.rs) on the batched GPU path."""

    def __init__(self, prop, kf_variant: KalmanVariant, sigma_reject: Optional[SigmaRejection], devices: Dict[str, GroundStation],
                 almanac: Optional[Almanac], msr_size: int = 2):
        self.prop = prop
        self.kf_variant = kf_variant
        self.sigma_reject = sigma_reject
        self.devices = dict(devices)
        self.almanac = almanac
        self.process_noise: List[ProcessNoise3D] = []
        self.max_step = 60 * NS_PER_S            # initializers.rs:71
        self.epoch_precision = 1_000             # 1 microsecond, initializers.rs:72
        self.msr_size = int(msr_size)

    new = classmethod(lambda cls, *a, **k: cls(*a, **k))

    def with_process_noise(self, snc: ProcessNoise3D) -> "KalmanODProcess":
        self.process_noise = [snc]
        return self

    # ---- packing
    def config_c(self) -> abi.OdConfigC:
        if len(self.process_noise) > 1:
            raise ODError("one process-noise model is supported on the GPU path")
        if self.max_step <= 0:
            raise ODError(f"StepSize: {self.max_step}")
        c = abi.OdConfigC()
        c.variant = int(self.kf_variant)
        c.msr_size = self.msr_size
        c.reject_num_sigmas = self.sigma_reject.num_sigmas if self.sigma_reject is not None else -1.0
        c.max_step_ns = int(self.max_step)
        c.epoch_precision_ns = int(self.epoch_precision)
        if self.process_noise:
            snc = self.process_noise[0]
            c.snc_enabled = 1
            c.snc_frame = 1 if snc.local_frame == LocalFrame.RIC else 0
            for i in range(3):
                c.snc_diag[i] = float(snc.diag[i])
            c.snc_disable_time_ns = int(snc.disable_time)
        return c

    def stations_c(self, frame: Frame):
        names = list(self.devices)
        arr = (abi.GroundStationC * max(len(names), 1))()
        for i, nme in enumerate(names):
            arr[i] = self.devices[nme].to_c(frame, self.almanac)
        return names, arr

    def process_arcs(self, initial_estimates: Sequence[KfEstimate], arc: TrackingDataArc, record_estimates: bool = False) -> ODSolution:
        """n independent `process_arc(initial_estimate_i, arc_i)` runs (od/process/mod.rs:128-497) in one launch."""
        n = len(initial_estimates)
        if arc.n != n:
            raise ODError(f"arc carries {arc.n} observation sets for {n} filters")
        if len(arc) < 2:
            raise ODError("TooFewMeasurements: need 2")  # process/mod.rs:139-145
        from .cosmic import pack_spacecraft

        frame = initial_estimates[0].nominal_state.orbit.frame
        st, cs, ep = pack_spacecraft(e.nominal_state for e in initial_estimates)
        cov0 = np.empty((81, n))
        for i, e in enumerate(initial_estimates):
            cov0[:, i] = np.asarray(e.covar, dtype=np.float64).reshape(9, 9).T.reshape(81)  # (c*9 + r)
        eng = self.prop.engine(frame, self.almanac)
        names, st_c = self.stations_c(frame)
        tracker = np.array([names.index(t) if t in names else -1 for t in arc.tracker], dtype=np.int32)
        res = eng.od_ekf_batch(self.config_c(), len(names), st_c, arc.epoch_ns, tracker, arc.obs, st, cs, ep, cov0,
                               record_estimates=record_estimates)
        res.templates = [e.nominal_state for e in initial_estimates]
        res.arc = arc
        res.devices = self.devices
        return res

    def process_arc(self, initial_estimate: KfEstimate, arc: TrackingDataArc) -> ODSolution:
        return self.process_arcs([initial_estimate], arc)


def SpacecraftKalmanOD(prop, kf_variant, sigma_reject, devices, almanac) -> KalmanODProcess:
    return KalmanODProcess(prop, kf_variant, sigma_reject, devices, almanac, msr_size=2)


def SpacecraftKalmanScalarOD(prop, kf_variant, sigma_reject, devices, almanac) -> KalmanODProcess:
    return KalmanODProcess(prop, kf_variant, sigma_reject, devices, almanac, msr_size=1)


# --------------------------------------------------------------------------- measurement simulation (host-side data generation)
def _rotation_matrix(rot, t_ns: int) -> np.ndarray:
    """inertial -> body-fixed DCM of the orientation model of include/nyxb.h (numpy restatement for the simulator)."""
    if rot is None or rot.kind == 0:
        return np.eye(3)
    d = duration_to_seconds(int(t_ns)) / 86400.0
    T = d / 36525.0
    ra = math.radians(rot.ra0_deg + rot.ra1_deg_cy * T)
    dec = math.radians(rot.dec0_deg + rot.dec1_deg_cy * T)
    w = math.radians(math.fmod(rot.w0_deg + rot.w1_deg_day * d, 360.0))
    sa, ca, sd, cd, sw, cw = math.sin(ra), math.cos(ra), math.sin(dec), math.cos(dec), math.sin(w), math.cos(w)
    ba = np.array([[-sa, ca, 0.0], [-sd * ca, -sd * sa, cd], [cd * ca, cd * sa, sd]])
    r3 = np.array([[cw, sw, 0.0], [-sw, cw, 0.0], [0.0, 0.0, 1.0]])
    return r3 @ ba


def station_state(gs: GroundStation, t_ns: int, integration_frame: Frame, almanac: Optional[Almanac]):
    """Inertial position/velocity of the antenna relative to the integration centre (trk_device.rs:150-152)."""
    pos_f, up = gs.body_fixed()
    R = _rotation_matrix(gs.frame.rotation, t_ns)
    wdot = math.radians(gs.frame.rotation.w1_deg_day) / 86400.0 if gs.frame.rotation is not None and gs.frame.rotation.kind else 0.0
    r = R.T @ pos_f
    v = R.T @ np.cross(np.array([0.0, 0.0, wdot]), pos_f)
    if gs.frame.ephemeris_id != integration_frame.ephemeris_id:
        b = almanac.bodies[almanac.body_index(gs.frame.ephemeris_id)]
        dt = 1_000_000_000
        p0 = b.position(t_ns)
        vb = (b.position(t_ns + dt) - b.position(t_ns - dt)) / (2.0 * dt / NS_PER_S)
        r, v = r + p0, v + vb
    return r, v, R.T @ up


def simulate_tracking(truth_epochs_ns, truth_states, devices: Dict[str, GroundStation], schedule: Sequence[str], frame: Frame,
                      almanac: Optional[Almanac], rng: Optional[np.random.Generator] = None) -> TrackingDataArc:
    """Synthetic range / Doppler observations of `truth_states[k]` ([m][6][n]) at `truth_epochs_ns[k]` from tracker
    `schedule[k]`; white noise of each station's sigma when `rng` is given.  Invisible passes are NaN (absent)."""
    truth_states = np.asarray(truth_states, dtype=np.float64)
    m, _, n = truth_states.shape
    obs = np.full((m, 2, n), np.nan)
    for k in range(m):
        gs = devices[schedule[k]]
        r_tx, v_tx, up_in = station_state(gs, int(truth_epochs_ns[k]), frame, almanac)
        rho = truth_states[k, :3, :] - r_tx[:, None]
        dv = truth_states[k, 3:6, :] - v_tx[:, None]
        rng_km = np.linalg.norm(rho, axis=0)
        rr = (rho * dv).sum(0) / rng_km
        elev = np.degrees(np.arcsin((rho * up_in[:, None]).sum(0) / rng_km))
        vis = elev >= gs.elevation_mask_deg
        if gs.frame.ephemeris_id != frame.ephemeris_id:
            # line of sight blocked by the body the spacecraft orbits (Vallado's SIGHT)
            r1, r2 = truth_states[k, :3, :], r_tx[:, None] * np.ones((1, n))
            r1sq, r2sq, r12 = (r1 * r1).sum(0), (r2 * r2).sum(0), (r1 * r2).sum(0)
            tau = (r1sq - r12) / (r1sq + r2sq - 2.0 * r12)
            blocked = (tau >= 0.0) & (tau <= 1.0) & ((1.0 - tau) * r1sq + r12 * tau <= frame.mean_equatorial_radius_km() ** 2)
            vis &= ~blocked
        for t in gs.measurement_types:
            val = rng_km if t == MeasurementType.Range else rr
            noise = rng.normal(0.0, gs.stochastic_noises[t].sigma, n) if rng is not None else 0.0
            obs[k, int(t), :] = np.where(vis, val + noise, np.nan)
    return TrackingDataArc(np.asarray(truth_epochs_ns, dtype=np.int64), list(schedule), obs)
