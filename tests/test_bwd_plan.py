"""Launch schedule of the fused backward (nans_clip_loss_bwd_plan): host-only invariants, no GPU.
148 SMs are assumed when no device is present, which is what a B200 has."""
import ctypes
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

from nans_clip_b200 import _lib  # noqa: E402

FIELDS = ["kind", "tile", "npass", "a_res", "nrb", "ntiles", "nr", "smem", "grid", "nsplit", "n_full", "ns_tail",
          "t1", "punits", "npairs", "zero"]
SMEM_CAP = 232448
PAIRS = 74


def plan(rows, N, D):
    lib = _lib.load()
    out = (ctypes.c_int64 * 16)()
    rc = lib.nans_clip_loss_bwd_plan(rows, N, D, ctypes.cast(out, ctypes.c_void_p), 16)
    assert rc == 0, lib.nans_last_error()
    return dict(zip(FIELDS, list(out)))


SHAPES = [(r, n, d) for d in (8, 64, 72, 256, 448, 512, 520, 640, 768, 776, 1024, 1032, 2048)
          for (r, n) in ((1, 1), (64, 64), (129, 1000), (300, 300), (4096, 4096), (4096, 32768), (8192, 32768),
                         (16384, 32768), (32768, 32768), (1024, 65536), (8192, 65536), (65536, 65536), (5000, 77777))]


@pytest.mark.parametrize("rows,N,D", SHAPES)
def test_schedule_invariants(rows, N, D, monkeypatch):
    for var in ("NANS_BWD_PERSIST", "NANS_BWD_NP", "NANS_BWD_1CTA", "NANS_NPP_T1"):
        monkeypatch.delenv(var, raising=False)
    p = plan(rows, N, D)
    assert 0 < p["smem"] <= SMEM_CAP and p["nr"] >= 2 and p["grid"] > 0
    assert p["ntiles"] == -(-N // p["tile"])
    if D <= 1024:
        assert p["kind"] in (2, 3)
        assert p["tile"] == (128 if 512 < D <= 768 else 256)
        assert p["a_res"] == (1 if D <= 768 else 0) and p["npass"] == (1 if D <= 768 else 2)
        assert p["nrb"] == -(-rows // 128)
    else:
        assert p["kind"] == 1 and p["tile"] == 128 and p["npass"] == -(-D // 256) and p["nrb"] == -(-rows // 256)
    units = 2 * p["nrb"] * p["npass"]
    if p["kind"] == 2:
        assert p["grid"] % 2 == 0 and p["n_full"] % PAIRS == 0 and 0 <= p["n_full"] <= units
        assert p["grid"] // 2 == p["n_full"] + (units - p["n_full"]) * p["ns_tail"]
        assert 1 <= p["ns_tail"] <= max(1, p["ntiles"])          # every split gets at least one tile
        assert p["zero"] == (0 if p["ns_tail"] == 1 else (2 if p["npass"] > 1 else 1))
        # the partial wave is what gets split: nothing is split when the units fill whole waves
        assert p["ns_tail"] == 1 or units - p["n_full"] > 0
    if p["kind"] == 3:
        assert D <= 512 and p["zero"] == 2 and p["grid"] == 2 * p["npairs"] and p["npairs"] <= PAIRS
        # helper schedule (the only persistent one)
        assert p["punits"] == units < p["npairs"] == PAIRS
        assert 1 <= p["t1"] < p["ntiles"]
        assert p["punits"] * (p["ntiles"] - p["t1"]) >= p["npairs"] - p["punits"]   # no idle helper
    if p["kind"] == 1:
        assert p["grid"] == 4 * p["nrb"] * p["npass"] * p["nsplit"] and p["zero"] == (2 if p["nsplit"] > 1 else 0)


def test_known_schedules(monkeypatch):
    for var in ("NANS_BWD_PERSIST", "NANS_BWD_NP", "NANS_BWD_1CTA", "NANS_NPP_T1"):
        monkeypatch.delenv(var, raising=False)
    # the bench size on one GPU: 512 units = 6 whole waves of 74 + 68 unsplit
    p = plan(32768, 32768, 512)
    assert (p["kind"], p["n_full"], p["ns_tail"], p["grid"], p["zero"]) == (2, 444, 1, 1024, 0)
    # 2 GPUs: 256 units = 3 waves + 34 units split in two
    p = plan(16384, 32768, 512)
    assert (p["kind"], p["n_full"], p["ns_tail"], p["grid"], p["zero"]) == (2, 222, 2, 2 * (222 + 68), 1)
    # 8 GPUs: 64 units on 74 pairs -> helper schedule, 112 of 128 tiles stay with the unit's own pair
    p = plan(4096, 32768, 512)
    assert (p["kind"], p["t1"], p["punits"], p["npairs"], p["zero"]) == (3, 112, 64, 74, 2)
    monkeypatch.setenv("NANS_BWD_PERSIST", "0")
    assert plan(4096, 32768, 512)["kind"] == 2
    monkeypatch.delenv("NANS_BWD_PERSIST")
    monkeypatch.setenv("NANS_BWD_NP", "0")
    assert plan(4096, 32768, 512)["kind"] == 1
    # the retired switches (single-CTA kernel, equal-range persistent schedule) no longer change anything
    monkeypatch.delenv("NANS_BWD_NP")
    monkeypatch.setenv("NANS_BWD_1CTA", "1")
    monkeypatch.setenv("NANS_BWD_PERSIST", "1")
    p = plan(4096, 32768, 512)
    assert (p["kind"], p["t1"]) == (3, 112)


def test_bad_arguments_are_rejected():
    lib = _lib.load()
    out = (ctypes.c_int64 * 16)()
    assert lib.nans_clip_loss_bwd_plan(0, 10, 64, ctypes.cast(out, ctypes.c_void_p), 16) < 0
    assert lib.nans_clip_loss_bwd_plan(10, 10, 12, ctypes.cast(out, ctypes.c_void_p), 16) < 0
    assert lib.nans_clip_loss_bwd_plan(10, 10, 64, ctypes.cast(out, ctypes.c_void_p), 4) < 0
    assert lib.nans_clip_loss_bwd_plan(10, 10, 64, None, 16) < 0
