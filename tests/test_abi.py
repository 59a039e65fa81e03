"""The C-ABI library loads and exports every symbol include/nyxb.h declares (no compute without a GPU)."""
import ctypes as C
import re
from pathlib import Path

import pytest

import nyx_b200 as nb
from nyx_b200 import abi

ROOT = Path(__file__).resolve().parent.parent


def test_header_symbols_exported():
    header = (ROOT / "include" / "nyxb.h").read_text()
    declared = set(re.findall(r"\b(nyxb_[a-z0-9_]+)\s*\(", header))
    lib = abi.load_library()
    assert declared == set(abi.EXPORTED_SYMBOLS), declared ^ set(abi.EXPORTED_SYMBOLS)
    for name in declared:
        assert hasattr(lib, name), name
    assert lib.nyxb_abi_version() == 4


def test_struct_layouts_match_header():
    assert C.sizeof(abi.IntegOpts) == 56   # ABI 4: + state_center (integration_frame)
    assert C.sizeof(abi.Rotation) == 56
    assert C.sizeof(abi.GravityFieldC) == 8 + 16 + 16 + 56 + 8   # ABI 4: + body
    assert C.sizeof(abi.BodyC) == 48
    assert C.sizeof(abi.SrpC) == 40
    assert C.sizeof(abi.DragC) == 40 + 56
    assert C.sizeof(abi.DynamicsC) == 64 + 32   # ABI 4: + point_mass_order[8]; n_gravity / n_point_masses reuse the pads
    assert C.sizeof(abi.Details) == 48


def test_product_fails_loudly_without_gpu():
    """No CPU fallback: without a CUDA device engine creation raises instead of computing on the host."""
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    prop = nb.Propagator.default(nb.SpacecraftDynamics.new(nb.OrbitalDynamics.two_body()))
    sc = nb.Spacecraft.from_orbit(nb.Orbit.cartesian(7000, 0, 0, 0, 7.5, 0, 0, nb.EARTH_J2000))
    with pytest.raises(nb.PropagationError, match="no CUDA device|CPU fallback"):
        prop.with_(sc).for_duration(60 * nb.Unit.Second)


def test_new_entry_points_reject_bad_arguments_and_need_a_gpu():
    """STM / OD / dispersion entry points: argument validation runs before any device work; without a CUDA device they
    report NYXB_RC_NO_DEVICE instead of computing on the host."""
    import numpy as np
    import torch

    lib = abi.load_library()
    vp = C.c_void_p
    assert lib.nyxb_propagate_batch_stm(None, 1, None, None, None, 0, None, None, None, None, None, None, None) == -1
    assert b"null" in lib.nyxb_last_error()
    assert lib.nyxb_od_ekf_batch(None, None, 0, None, None, 1, None, None, None, None, None) == -1
    assert lib.nyxb_mvn_sample(0, 1, 0, 4, None, None, None, None, None) == -1
    if not torch.cuda.is_available():
        t = np.zeros(9); L = np.eye(9).reshape(81); out = np.zeros((9, 4))
        rc = lib.nyxb_mvn_sample(0, 1, 0, 4, t.ctypes.data, None, L.ctypes.data, out.ctypes.data, None)
        assert rc == -2 and b"no CUDA device" in lib.nyxb_last_error()   # NYXB_RC_NO_DEVICE: no CPU fallback


def test_od_struct_layouts():
    assert C.sizeof(abi.GroundStationC) == 176
    assert C.sizeof(abi.OdConfigC) == 72
    assert C.sizeof(abi.TrackingArcC) == 32
    assert C.sizeof(abi.OdOutputsC) == 96


def test_missing_library_is_an_error(monkeypatch, tmp_path):
    monkeypatch.setattr(abi, "_lib", None)
    monkeypatch.setenv("NYXB_LIBRARY", str(tmp_path / "nope.so"))
    with pytest.raises(nb.NyxbLibraryMissing):
        abi.load_library()
    monkeypatch.delenv("NYXB_LIBRARY")
    monkeypatch.setattr(abi, "_lib", None)


def test_product_does_not_import_oracle():
    for py in (ROOT / "nyx_b200").rglob("*.py"):
        src = py.read_text()
        assert "oracle" not in src.replace("CPU oracle", "").replace("the oracle", "").lower() or "import" not in [
            l for l in src.splitlines() if "oracle" in l.lower() and ("import" in l)], py
    for cu in (ROOT / "nyx_b200" / "csrc").glob("*.cu*"):
        assert "#include \"../../oracle" not in cu.read_text() and "nyx_oracle.h" not in cu.read_text(), cu
