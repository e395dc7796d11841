"""Binary feature shards: host-side format logic, no GPU."""
import json
import os
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

from nans_clip_b200.eval import feature_io as fio  # noqa: E402


def _rows(n, d, seed=0):
    rng = np.random.default_rng(seed)
    return np.arange(n, dtype=np.int64) * 3 + 1000000, rng.standard_normal((n, d)).astype(np.float32)


@pytest.mark.parametrize("n,d", [(0, 8), (1, 8), (7, 16), (1000, 72)])
@pytest.mark.parametrize("with16", [False, True])
def test_round_trip(tmp_path, n, d, with16):
    ids, f32 = _rows(n, d)
    f16 = f32.astype(np.float16).view(np.uint16) if with16 else None
    p = str(tmp_path / "a.nansf")
    fio.write_shard(p, ids, f32, f16, fio.DT16_F16 if with16 else fio.DT16_NONE, normalized=True)
    assert fio.is_shard(p)
    for mmap in (True, False):
        i2, a32, a16, h = fio.read_shard(p, mmap=mmap)
        assert h == {"rows": n, "D": d, "dtype16": fio.DT16_F16 if with16 else 0, "normalized": True}
        assert np.array_equal(np.asarray(i2), ids) and np.array_equal(np.asarray(a32), f32)
        assert (a16 is None) == (not with16)
        if with16:
            assert np.array_equal(np.asarray(a16), f16)
    # every section is 64-byte aligned and the file has no tail
    off_ids, off_f32, off_f16, end = fio._layout(n, d, fio.DT16_F16 if with16 else 0)
    assert off_ids % 64 == 0 and off_f32 % 64 == 0 and off_f16 % 64 == 0 and os.path.getsize(p) == end


def test_jsonl_compat_and_conversion(tmp_path):
    ids, f32 = _rows(33, 24, 1)
    j = tmp_path / "f.jsonl"
    with open(j, "w") as f:
        for i, r in zip(ids.tolist(), f32.tolist()):
            f.write(json.dumps({"image_id": i, "feature": r}) + "\n")
        f.write("\n")  # blank lines are skipped like the reference's strip()
    a_ids, a32, a16, code = fio.load_features(str(j), "image_id")
    assert a_ids == ids.tolist() and np.allclose(a32, f32, atol=0) and a16 is None and code == 0
    s = tmp_path / "f.nansf"
    assert fio.jsonl_to_shard(str(j), "image_id", str(s)) == 33
    b_ids, b32, b16, code = fio.load_features(str(s), "image_id")
    assert np.array_equal(np.asarray(b_ids), ids) and np.array_equal(np.asarray(b32), a32) and b16 is None
    assert not fio.is_shard(str(j)) and not fio.is_shard(str(tmp_path / "missing"))


def test_bad_files_are_rejected(tmp_path):
    ids, f32 = _rows(10, 8)
    p = str(tmp_path / "a.nansf")
    fio.write_shard(p, ids, f32)
    raw = open(p, "rb").read()
    open(tmp_path / "trunc.nansf", "wb").write(raw[:-16])
    with pytest.raises(ValueError, match="truncated"):
        fio.read_shard(str(tmp_path / "trunc.nansf"))
    open(tmp_path / "ver.nansf", "wb").write(raw[:8] + (99).to_bytes(4, "little") + raw[12:])
    with pytest.raises(ValueError, match="version"):
        fio.read_shard(str(tmp_path / "ver.nansf"))
    open(tmp_path / "magic.nansf", "wb").write(b"XXXXXXXX" + raw[8:])
    with pytest.raises(ValueError, match="magic"):
        fio.read_header(str(tmp_path / "magic.nansf"))
    with pytest.raises(ValueError):
        fio.write_shard(p, ids[:5], f32)                       # row mismatch
    with pytest.raises(ValueError):
        fio.write_shard(p, ids, f32, f32.astype(np.float16).view(np.uint16))  # feat16 without its dtype


def test_shard_slices_are_rank_shards(tmp_path):
    """A contiguous row range of the mmap is what a rank of the sharded retrieval loads."""
    ids, f32 = _rows(101, 16, 2)
    p = str(tmp_path / "g.nansf")
    fio.write_shard(p, ids, f32, f32.astype(np.float16).view(np.uint16), fio.DT16_F16)
    _, a32, a16, _ = fio.read_shard(p)
    W = 4
    parts = [np.ascontiguousarray(a32[101 * r // W:101 * (r + 1) // W]) for r in range(W)]
    assert np.array_equal(np.concatenate(parts), f32)
    assert np.array_equal(np.concatenate([np.asarray(a16[101 * r // W:101 * (r + 1) // W]) for r in range(W)]).view(np.float16),
                          f32.astype(np.float16))


# ---- native JSONL parser (csrc/jsonl.cu) ----------------------------------------------------------
def _write(path, lines):
    with open(path, "w", newline="") as f:
        f.write("".join(lines))


def test_native_jsonl_parser_is_bit_identical_to_json_loads(tmp_path):
    rng = np.random.default_rng(3)
    n, d = 500, 96
    f32 = (rng.standard_normal((n, d)) * np.exp(rng.uniform(-12, 3, (n, 1)))).astype(np.float32)
    ids = rng.integers(-5, 2 ** 40, n)
    p = tmp_path / "f.jsonl"
    _write(p, [json.dumps({"image_id": int(i), "feature": r.tolist()}) + "\n" for i, r in zip(ids, f32)])
    a_ids, a = fio._load_jsonl_python(str(p), "image_id")
    for threads in (1, 3, 0):
        b_ids, b = fio._load_jsonl_native(str(p), "image_id", threads)
        assert a_ids == b_ids and b.dtype == np.float32 and np.array_equal(a.view(np.uint32), b.view(np.uint32))
    c_ids, c = fio.load_jsonl_features(str(p), "image_id")
    assert c_ids == a_ids and np.array_equal(c, a)


def test_native_jsonl_parser_formats(tmp_path):
    """Exponents, integers, negative zero, doubles that round to float32, key order, blank lines,
    CRLF, extra keys, missing final newline."""
    lines = ['{"text_id": 7, "feature": [1, -0.0, 2.5e-05, 1E+2, 0.1, 16777217, 3.4028234663852886e+38]}\n',
             '\n',
             '{"feature": [ 0.30000000000000004 ,1e-46, -1.5,4,5 ,6, 1.0000000596046448], "text_id":8 }\r\n',
             '   \n',
             '{"extra": "x\\"y", "text_id": 9, "feature": [0,0,0,0,0,0,0], "more": [1, 2]}']
    p = tmp_path / "g.jsonl"
    _write(p, lines)
    a_ids, a = fio._load_jsonl_python(str(p), "text_id")
    b_ids, b = fio._load_jsonl_native(str(p), "text_id")
    assert a_ids == b_ids == [7, 8, 9]
    assert np.array_equal(a.view(np.uint32), b.view(np.uint32))      # incl. the sign of -0.0


def test_native_jsonl_parser_falls_back(tmp_path):
    """What the native parser refuses still loads through json.loads, like the reference would."""
    p = tmp_path / "s.jsonl"
    _write(p, ['{"image_id": "a7", "feature": [1.0, 2.0]}\n', '{"image_id": "b8", "feature": [3.0, 4.0]}\n'])
    with pytest.raises(Exception):
        fio._load_jsonl_native(str(p), "image_id")
    ids, f = fio.load_jsonl_features(str(p), "image_id")
    assert ids == ["a7", "b8"] and np.array_equal(f, np.array([[1, 2], [3, 4]], dtype=np.float32))
    r = tmp_path / "r.jsonl"                                        # ragged rows
    _write(r, ['{"image_id": 1, "feature": [1.0, 2.0]}\n', '{"image_id": 2, "feature": [3.0]}\n'])
    with pytest.raises(Exception):
        fio._load_jsonl_native(str(r), "image_id")
    t = tmp_path / "t.jsonl"                                        # truncated last line
    _write(t, ['{"image_id": 1, "feature": [1.0, 2.0]}\n', '{"image_id": 2, "feature": [3.0, 4'])
    with pytest.raises(Exception):
        fio._load_jsonl_native(str(t), "image_id")
    # tokens from_chars would take but json.loads reads differently (or refuses): the native parser must
    # refuse them so that the fallback decides
    for k, bad in enumerate(['{"image_id": 12.5, "feature": [1.0]}\n', '{"image_id": 1e3, "feature": [1.0]}\n',
                             '{"image_id": 1, "feature": [nan]}\n', '{"image_id": 1, "feature": [inf]}\n',
                             '{"image_id": 1, "feature": [012]}\n', '{"image_id": 1, "feature": [.5]}\n',
                             '{"image_id": 1, "feature": [5.]}\n', '{"image_id": 1, "feature": [1.0]} trailing\n',
                             '{"image_id": 1, "feature": [1.0]\n', '"image_id": 1, "feature": [1.0]}\n']):
        b = tmp_path / f"bad{k}.jsonl"
        _write(b, [bad])
        with pytest.raises(Exception):
            fio._load_jsonl_native(str(b), "image_id")
    nn = tmp_path / "nan.jsonl"                                     # json.loads' own NaN / Infinity spelling
    _write(nn, ['{"image_id": 1, "feature": [NaN, -Infinity]}\n'])
    ids, f = fio.load_jsonl_features(str(nn), "image_id")
    assert ids == [1] and np.isnan(f[0, 0]) and f[0, 1] == -np.inf
    fl = tmp_path / "fl.jsonl"                                      # a float id stays a float, as json.loads has it
    _write(fl, ['{"image_id": 12.5, "feature": [1.0]}\n'])
    ids, f = fio.load_jsonl_features(str(fl), "image_id")
    assert ids == [12.5]
    e = tmp_path / "e.jsonl"
    _write(e, [])
    ids, f = fio.load_jsonl_features(str(e), "image_id")
    assert ids == [] and f.shape[0] == 0


def test_sharded_jsonl_load_parses_only_the_rank_rows(tmp_path):
    """load_features(shard=(rank, world)): ids of the whole file, feature rows of the rank's contiguous
    share only (what a torchrun'd make_topk_predictions does per rank), native and json.loads paths."""
    rng = np.random.default_rng(5)
    n, d, W = 103, 24, 4
    f32 = rng.standard_normal((n, d)).astype(np.float32)
    ids = rng.integers(0, 10 ** 9, n).tolist()
    p = tmp_path / "g.jsonl"
    _write(p, [json.dumps({"image_id": i, "feature": r.tolist()}) + "\n" for i, r in zip(ids, f32)])
    for native in (True, False):
        got_rows = []
        for r in range(W):
            if native:
                a_ids, a = fio._load_jsonl_native(str(p), "image_id", shard=(r, W))
            else:
                os.environ["NANS_JSONL_PYTHON"] = "1"
                try:
                    a_ids, a, _, _ = fio.load_features(str(p), "image_id", shard=(r, W))
                finally:
                    del os.environ["NANS_JSONL_PYTHON"]
            assert a_ids == ids and a.shape == (n * (r + 1) // W - n * r // W, d)
            got_rows.append(a)
        assert np.array_equal(np.concatenate(got_rows), f32)
    s = tmp_path / "g.nansf"
    fio.jsonl_to_shard(str(p), "image_id", str(s))
    parts = [np.array(fio.load_features(str(s), "image_id", shard=(r, W))[1]) for r in range(W)]
    assert np.array_equal(np.concatenate(parts), f32)
    # a malformed feature list outside the rank's rows does not concern that rank; inside it does
    bad = tmp_path / "bad.jsonl"
    lines = [json.dumps({"image_id": i, "feature": r.tolist()}) + "\n" for i, r in zip(ids, f32)]
    lines[1] = '{"image_id": 5, "feature": [oops]}\n'
    _write(bad, lines)
    fio._load_jsonl_native(str(bad), "image_id", shard=(3, W))
    with pytest.raises(Exception):
        fio._load_jsonl_native(str(bad), "image_id", shard=(0, W))
