"""Gravity-field coefficient loading — host-side mirror of ``io/gravity.rs``.

``GravityFieldData.{from_cof,from_shadr,from_j2}`` follow the reference loaders'
semantics (io/gravity.rs:117-128, 150-367, 370-501): coefficients are already
normalised in the files, rows beyond the requested degree stop the scan, orders
beyond the requested order are skipped, and the resulting ``degree``/``order`` are
the maxima *seen in the file* (not the requested ones).

The packed fixtures under ``data/`` (``*.npz``) are produced by
``scripts/make_gravity_fixtures.py`` from the public JGM-3 / GRAIL files so that
nothing needs ``/root/reference`` at run time.
"""
from __future__ import annotations

import gzip
from dataclasses import dataclass
from pathlib import Path

import numpy as np

from .frames import Frame

_DATA_DIR = Path(__file__).resolve().parent.parent / "data"


def _split_cof_pair(item: str):
    """COF quirk (io/gravity.rs:236-312): C and S are glued together when S < 0."""
    n_minus = item.count("-")
    if (n_minus == 3 and not item.startswith("-")) or n_minus == 4:
        parts = item.split("-")
        if len(parts) == 5:  # both negative
            return float("-" + parts[1] + "-" + parts[2]), float("-" + parts[3] + "-" + parts[4])
        return float(parts[0] + "-" + parts[1]), float("-" + parts[2] + "-" + parts[3])
    return float(item), None


@dataclass
class GravityFieldData:
    """``GravityFieldData`` (io/gravity.rs:90-96): normalised C̄nm, S̄nm + the body-fixed frame."""

    degree: int
    order: int
    c_nm: np.ndarray  # [(degree+1), (degree+1)] row-major, (n, m)
    s_nm: np.ndarray
    frame: Frame

    def max_degree_n(self) -> int:
        return self.degree

    def max_order_m(self) -> int:
        return self.order

    def cs_nm(self, degree: int, order: int):
        return float(self.c_nm[degree, order]), float(self.s_nm[degree, order])

    # ------------------------------------------------------------------ constructors
    @classmethod
    def from_j2(cls, j2: float, frame: Frame) -> "GravityFieldData":
        """io/gravity.rs:117-128 — the value is stored as-is at C̄[2][0]."""
        c = np.zeros((3, 3))
        c[2, 0] = j2
        return cls(2, 0, c, np.zeros((3, 3)), frame)

    @classmethod
    def _from_lines(cls, rows, degree, order, frame):
        c = np.zeros((degree + 1, degree + 1))
        s = np.zeros((degree + 1, degree + 1))
        max_deg = max_ord = 0
        for n, m, cv, sv in rows:
            if n > degree:
                break  # file is organised by degree (io/gravity.rs:335-339)
            if m <= order:
                c[n, m] = cv
                s[n, m] = sv
            max_ord = max(max_ord, m)
            max_deg = max(max_deg, n)
        return cls(max_deg, max_ord, c, s, frame)

    @classmethod
    def from_cof(cls, filepath, degree: int, order: int, gunzipped: bool, frame: Frame):
        """io/gravity.rs:150-367."""
        raw = gzip.open(filepath, "rt").read() if gunzipped else Path(filepath).read_text()

        def rows():
            for line in raw.split("\n"):
                if not line or not line.startswith("R"):
                    continue
                items = line.split()
                n, m = int(items[1]), int(items[2])
                cv, sv = 0.0, 0.0
                if len(items) > 3:
                    if degree == 0:
                        cv = float(items[3])
                    else:
                        cv, glued = _split_cof_pair(items[3])
                        if glued is not None:
                            sv = glued
                if len(items) > 4:
                    sv = float(items[4])
                yield n, m, cv, sv

        return cls._from_lines(rows(), degree, order, frame)

    @classmethod
    def from_shadr(cls, filepath, degree: int, order: int, gunzipped: bool, frame: Frame):
        """io/gravity.rs:139-147 + 370-501 (first line is a header and is skipped)."""
        opener = gzip.open if gunzipped else open

        def rows():
            with opener(filepath, "rt") as fh:
                for lno, line in enumerate(fh):
                    if lno == 0:
                        continue
                    items = line.replace(",", " ").split()
                    if len(items) < 2:
                        yield 0, 0, 0.0, 0.0
                        continue
                    n, m = int(items[0]), int(items[1])
                    cv = float(items[2].replace("D", "E")) if len(items) > 2 else 0.0
                    sv = float(items[3].replace("D", "E")) if len(items) > 3 else 0.0
                    yield n, m, cv, sv

        return cls._from_lines(rows(), degree, order, frame)

    @classmethod
    def from_fixture(cls, name: str, degree: int, order: int, frame: Frame):
        """Load a packed fixture (``data/<name>.npz``: arrays n, m, c, s in file order) with the
        same truncation semantics as from_cof/from_shadr."""
        z = np.load(_DATA_DIR / f"{name}.npz")
        rows = zip(z["n"].tolist(), z["m"].tolist(), z["c"].tolist(), z["s"].tolist())
        return cls._from_lines(rows, degree, order, frame)
