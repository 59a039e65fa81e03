"""`Traj` — recorded trajectory of one spacecraft with Hermite interpolation.

Mirrors ``md/trajectory/traj.rs:54-162`` for what the propagation path produces:
``finalize`` (dedup equal epochs, sort by epoch, traj.rs:75-80), ``at`` (exact hit or a 13-sample window
centred like the reference's, traj.rs:83-126) and the `Interpolatable for Spacecraft` rule
(interpolatable.rs:53-108): position/velocity by Hermite interpolation of (r, v) pairs.

anise's `hermite_eval` is not in the reference tree; the interpolation here is the textbook Hermite
divided-difference form (the algorithm NAIF's HRMINT / SPK type 13 uses) — parity unpinned at the anise boundary.
Host-side utility (numpy): the reference interpolates on the host as well; it is not on the hot path.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Optional

import numpy as np

from .cosmic import Spacecraft

INTERPOLATION_SAMPLES = 13  # interpolatable.rs:22


class TrajError(RuntimeError):
    """`TrajError::NoInterpolationData` (md/trajectory/mod.rs)."""


def hermite_eval(xs: np.ndarray, ys: np.ndarray, ydots: np.ndarray, x: float):
    """Hermite interpolation through (xs, ys) with derivatives ydots; returns (y(x), y'(x))."""
    n = len(xs)
    z = np.repeat(np.asarray(xs, dtype=np.float64), 2)
    q = np.zeros((2 * n, 2 * n))
    q[0::2, 0] = ys
    q[1::2, 0] = ys
    q[1::2, 1] = ydots
    q[2::2, 1] = (q[2::2, 0] - q[1:-1:2, 0]) / (z[2::2] - z[1:-1:2])
    for j in range(2, 2 * n):
        q[j:, j] = (q[j:, j - 1] - q[j - 1:-1, j - 1]) / (z[j:] - z[:-j])
    coef = np.diag(q)
    # Horner on the Newton form, value and derivative together
    val, der = coef[-1], 0.0
    for k in range(2 * n - 2, -1, -1):
        der = der * (x - z[k]) + val
        val = val * (x - z[k]) + coef[k]
    return float(val), float(der)


@dataclass
class Traj:
    """Recorded states of one spacecraft: epochs [k] (int64 ns) and states [k, 6] (km, km/s)."""

    template: Spacecraft
    epochs_ns: np.ndarray
    states: np.ndarray
    name: Optional[str] = None

    def finalize(self) -> "Traj":
        """traj.rs:75-80: remove duplicate epochs, sort by epoch (a back-propagation is stored ascending)."""
        keep = np.ones(len(self.epochs_ns), dtype=bool)
        keep[1:] = self.epochs_ns[1:] != self.epochs_ns[:-1]
        ep, st = self.epochs_ns[keep], self.states[keep]
        order = np.argsort(ep, kind="stable")
        self.epochs_ns, self.states = ep[order], st[order]
        return self

    def __len__(self) -> int:
        return len(self.epochs_ns)

    def _sc(self, epoch_ns: int, rv) -> Spacecraft:
        vec = self.template.to_vector()
        vec[:6] = rv
        return self.template.with_vector(int(epoch_ns), vec)

    def first(self) -> Spacecraft:
        return self._sc(self.epochs_ns[0], self.states[0])

    def last(self) -> Spacecraft:
        return self._sc(self.epochs_ns[-1], self.states[-1])

    def at(self, epoch_ns: int) -> Spacecraft:
        """traj.rs:83-126."""
        n = len(self)
        if n == 0 or self.epochs_ns[0] > epoch_ns or self.epochs_ns[-1] < epoch_ns:
            raise TrajError(f"no interpolation data at {epoch_ns}")
        idx = int(np.searchsorted(self.epochs_ns, epoch_ns, side="left"))
        if idx < n and self.epochs_ns[idx] == epoch_ns:
            return self._sc(epoch_ns, self.states[idx])  # "Oh wow, we actually had this exact state!"
        if idx == 0 or idx >= n:
            raise TrajError(f"no interpolation data at {epoch_ns}")
        num_left = INTERPOLATION_SAMPLES // 2
        first_idx = max(idx - num_left, 0)
        last_idx = min(n, first_idx + INTERPOLATION_SAMPLES)
        if last_idx == n:
            first_idx = max(last_idx - 2 * num_left, 0)
        t0 = int(self.epochs_ns[first_idx])
        ts = (self.epochs_ns[first_idx:last_idx] - t0).astype(np.float64) * 1e-9
        win = self.states[first_idx:last_idx]
        x = (int(epoch_ns) - t0) * 1e-9
        rv = np.empty(6)
        for c in range(3):
            rv[c], rv[3 + c] = hermite_eval(ts, win[:, c], win[:, 3 + c], x)
        return self._sc(epoch_ns, rv)

    # ---- iteration and export (traj.rs:148-193, 226-360; traj_it.rs:32-63)
    def every_between(self, step_ns: int, start_ns: int, end_ns: int):
        """`Traj::every_between`: states every `step` over TimeSeries::inclusive(max(start, first), min(end, last), step);
        the iteration ends at the first epoch without interpolation data."""
        step_ns = int(step_ns)
        if step_ns <= 0:
            raise ValueError("step must be positive")
        if len(self) == 0:
            return
        t = max(int(start_ns), int(self.epochs_ns[0]))
        end = min(int(end_ns), int(self.epochs_ns[-1]))
        while t <= end:
            try:
                yield self.at(t)
            except TrajError:
                return
            t += step_ns

    def every(self, step_ns: int):
        """`Traj::every` (traj.rs:148-150)."""
        if len(self) == 0:
            return iter(())
        return self.every_between(step_ns, int(self.epochs_ns[0]), int(self.epochs_ns[-1]))

    def filter_by_epoch(self, start_ns: int, end_ns: int) -> "Traj":
        """`Traj::filter_by_epoch` with an inclusive range (traj.rs:165-193): only the recorded states inside it."""
        keep = (self.epochs_ns >= int(start_ns)) & (self.epochs_ns <= int(end_ns))
        return Traj(self.template, self.epochs_ns[keep].copy(), self.states[keep].copy(), self.name)

    def to_parquet(self, path, fields=None, start_ns: Optional[int] = None, end_ns: Optional[int] = None,
                   step_ns: Optional[int] = None, metadata: Optional[dict] = None):
        """`Traj::to_parquet` (traj.rs:226-360): "Epoch (UTC)" + the requested fields (default `Spacecraft::export_params`) of all
        recorded states, or of the states interpolated every `step` (default 1 min) when start/end/step is given; fields no
        state provides are dropped.  (For whole ensembles use `Results.to_parquet`, which interpolates on the device.)"""
        import pyarrow as pa
        import pyarrow.parquet as pq

        from .cosmic import epochs_to_utc_iso
        from .param import EXPORT_PARAMS, StateError, evaluate

        if start_ns is not None or end_ns is not None or step_ns is not None:
            sts = list(self.every_between(60 * 10**9 if step_ns is None else step_ns,
                                          int(self.epochs_ns[0]) if start_ns is None else start_ns,
                                          int(self.epochs_ns[-1]) if end_ns is None else end_ns))
            epochs = np.array([s.epoch() for s in sts], dtype=np.int64)
            rv = np.array([s.orbit.to_cartesian_pos_vel() for s in sts]).reshape(len(sts), 6).T
        else:
            epochs, rv = self.epochs_ns, self.states.T
        frame = self.template.orbit.frame
        cols = [pa.array(epochs_to_utc_iso(epochs), type=pa.string())]
        schema = [pa.field("Epoch (UTC)", pa.string(), nullable=False)]
        for f in (EXPORT_PARAMS if fields is None else fields):
            try:
                vals = evaluate(f, rv, frame.mu_km3_s2(), self.template)
            except StateError:
                continue
            cols.append(pa.array(vals, type=pa.float64()))
            schema.append(pa.field(str(f), pa.float64(), nullable=False, metadata={"unit": f.unit, "Frame": frame.name}))
        meta = {"Purpose": "Trajectory data"}
        meta.update(metadata or {})
        pq.write_table(pa.Table.from_arrays(cols, schema=pa.schema(schema, metadata=meta)), str(path))
        return path

    @classmethod
    def from_parquet(cls, path, template: Spacecraft) -> "Traj":
        """`Traj::<Spacecraft>::from_parquet` (md/trajectory/sc_traj.rs:212-440): "Epoch (UTC)" and the six Cartesian columns are
        required; dry / prop mass columns update the template when present.  The frame is the template's: the reference reads it
        from the Dhall-serialised field metadata, here the field metadata carries the frame's name and is checked against it."""
        import pyarrow.parquet as pq

        from .cosmic import utc_iso_to_epochs
        from .param import StateParameter as P

        tab = pq.read_table(str(path))
        names = set(tab.column_names)
        if "Epoch (UTC)" not in names:
            raise TrajError("MissingData: Epoch (UTC)")
        cart = [P.X, P.Y, P.Z, P.VX, P.VY, P.VZ]
        for f in cart:
            if str(f) not in names:
                raise TrajError(f"MissingData: {f}")
        meta = tab.schema.field(str(P.X)).metadata or {}
        frame_name = meta.get(b"Frame")
        if frame_name is None:
            raise TrajError("MissingData: Frame in metadata")
        if frame_name.decode() != template.orbit.frame.name:
            raise TrajError(f"trajectory is in frame {frame_name.decode()}, the template in {template.orbit.frame.name}")
        ep = utc_iso_to_epochs(tab["Epoch (UTC)"].to_pylist())
        states = np.column_stack([np.asarray(tab[str(f)].to_pylist(), dtype=np.float64) for f in cart])
        return cls(template, ep, states).finalize()

    # ---- CCSDS OEM (KVN) — md/trajectory/sc_traj.rs:176-210; the reference goes through anise's `Ephemeris`, which is not in the
    # tree: read / written here from the published layout of CCSDS 502.0-B (header, META_START..META_STOP, ephemeris lines).
    @classmethod
    def from_oem_file(cls, path, template: Optional[Spacecraft] = None) -> "Traj":
        """`Traj::from_oem_file`: all ephemeris lines of all segments; duplicate epochs removed (`finalize`); the trajectory is
        named after OBJECT_ID.  Without a template a massless default spacecraft is used, in the frame the metadata names
        (CENTER_NAME Earth / Moon with an inertial REF_FRAME)."""
        from .cosmic import Orbit, utc_iso_to_epochs
        from .frames import EARTH_J2000, MOON_J2000, NS_PER_S

        meta, rows, in_meta, in_cov = {}, [], False, False
        for raw in open(path, "r"):
            line = raw.strip()
            if not line or line.startswith("COMMENT"):
                continue
            if line == "META_START":
                in_meta = True
            elif line == "META_STOP":
                in_meta = False
            elif line == "COVARIANCE_START":
                in_cov = True
            elif line == "COVARIANCE_STOP":
                in_cov = False
            elif in_cov:
                continue
            elif "=" in line:
                key, val = (t.strip() for t in line.split("=", 1))
                if in_meta or key in ("CCSDS_OEM_VERS", "CREATION_DATE", "ORIGINATOR"):
                    meta.setdefault(key, val)
            else:
                tok = line.split()
                if len(tok) < 7:
                    raise TrajError(f"malformed OEM ephemeris line: {line!r}")
                rows.append((tok[0], [float(v) for v in tok[1:7]]))
        if "CCSDS_OEM_VERS" not in meta or not rows:
            raise TrajError("not a CCSDS OEM file (no version keyword or no ephemeris data)")
        scale = meta.get("TIME_SYSTEM", "UTC").upper()
        stamps = [r[0] for r in rows]
        if scale == "UTC":
            ep = utc_iso_to_epochs(stamps)
        else:
            offset = {"TT": 0, "TDB": 0, "TAI": 32_184_000_000, "GPS": 19 * NS_PER_S + 32_184_000_000}.get(scale)
            if offset is None:
                raise TrajError(f"unsupported OEM TIME_SYSTEM {scale}")
            ep = (np.array(stamps, dtype="datetime64[ns]") - np.datetime64("2000-01-01T12:00:00", "ns")).astype(np.int64) + offset
        if template is None:
            center = meta.get("CENTER_NAME", "EARTH").upper()
            if meta.get("REF_FRAME", "ICRF").upper() not in ("ICRF", "EME2000", "J2000", "GCRF"):
                raise TrajError(f"unsupported OEM REF_FRAME {meta.get('REF_FRAME')}")
            frame = {"EARTH": EARTH_J2000, "MOON": MOON_J2000}.get(center)
            if frame is None:
                raise TrajError(f"unsupported OEM CENTER_NAME {center}")
            template = Spacecraft.from_orbit(Orbit.cartesian(*rows[0][1], int(ep[0]), frame))
        tr = cls(template, np.asarray(ep, dtype=np.int64), np.array([r[1] for r in rows]), meta.get("OBJECT_ID"))
        order = np.argsort(tr.epochs_ns, kind="stable")     # finalize() dedups neighbours: sort first
        tr.epochs_ns, tr.states = tr.epochs_ns[order], tr.states[order]
        return tr.finalize()

    def to_oem_file(self, path, object_id: str, originator: Optional[str] = None, object_name: Optional[str] = None,
                    start_ns: Optional[int] = None, end_ns: Optional[int] = None, step_ns: Optional[int] = None):
        """`Traj::to_oem_file`: the recorded states, or (when start / end / step is given) the states interpolated every `step`
        (default 1 min) like `to_parquet`; UTC time stamps with nanosecond digits, values with full precision."""
        from .cosmic import epochs_to_utc_iso

        if start_ns is not None or end_ns is not None or step_ns is not None:
            sts = list(self.every_between(60 * 10**9 if step_ns is None else step_ns,
                                          int(self.epochs_ns[0]) if start_ns is None else start_ns,
                                          int(self.epochs_ns[-1]) if end_ns is None else end_ns))
            epochs = np.array([s.epoch() for s in sts], dtype=np.int64)
            rv = np.array([s.orbit.to_cartesian_pos_vel() for s in sts]).reshape(len(sts), 6)
        else:
            epochs, rv = self.epochs_ns, self.states
        if len(epochs) == 0:
            raise TrajError("no state to export")
        iso = epochs_to_utc_iso(epochs)
        frame = self.template.orbit.frame
        center = {399: "EARTH", 301: "MOON", 10: "SUN"}.get(frame.ephemeris_id, str(frame.ephemeris_id))
        with open(path, "w") as fh:
            fh.write("CCSDS_OEM_VERS = 2.0\n")
            fh.write(f"CREATION_DATE = {np.datetime_as_string(np.datetime64('now'), unit='s')}\n")
            fh.write(f"ORIGINATOR = {originator or 'nyx_b200'}\n\nMETA_START\n")
            fh.write(f"OBJECT_NAME = {object_name or object_id}\nOBJECT_ID = {object_id}\nCENTER_NAME = {center}\nREF_FRAME = ICRF\n")
            fh.write(f"TIME_SYSTEM = UTC\nSTART_TIME = {iso[0]}\nSTOP_TIME = {iso[-1]}\nMETA_STOP\n\n")
            for t, row in zip(iso, rv):
                fh.write(t + " " + " ".join(repr(float(v)) for v in row) + "\n")
        return path
