# -*- coding: utf-8 -*-
"""Binary feature shards for retrieval.

The reference stores features as JSON text — extract_features.py:179-181 / 200-202 write
`{"image_id": int, "feature": [floats]}` per line, make_topk_predictions.py:57-65 parses them back
at a few thousand lines per second, so a 1M-row gallery costs minutes of parsing in front of a
millisecond kernel.  This module adds a binary shard that keeps exactly the same content:

    offset 0    magic  b"NANSFEAT"              8 bytes
           8    version (=1)                     uint32
          12    D                                uint32
          16    rows                             uint64
          24    dtype16: 0 none, 1 fp16, 2 bf16  uint32    (same codes as NANS_F16 / NANS_BF16)
          28    normalized flag                  uint32
          32    reserved                         32 bytes
          64    ids     int64  [rows]            (padded to 64 bytes)
                feat32  fp32   [rows, D]         (padded to 64 bytes)   — what the fp32 rescoring reads
                feat16  16-bit [rows, D]         (if dtype16 != 0)      — the tensor-core operand copy

Every section starts 64-byte aligned, so the arrays are `np.memmap`-able and a contiguous row range
(one rank's gallery shard) is one slice.  JSONL stays the compatibility format: `load_features`
reads either (it sniffs the magic), `jsonl_to_shard` converts, `FeatureWriter.append` is the
"normalise -> fp32 + 16-bit shard" writer for an extract_features-style loop (it runs kernel (1)
on the device, i.e. the `features /= features.norm(dim=-1, keepdim=True)` of extract_features.py:178
fused with the 16-bit cast).
"""
from __future__ import annotations

import json
import os
import struct
from typing import Optional

import numpy as np

MAGIC = b"NANSFEAT"
VERSION = 1
HEADER_BYTES = 64
_HDR = struct.Struct("<8sIIQII32x")
DT16_NONE, DT16_F16, DT16_BF16 = 0, 1, 2


def _pad64(n: int) -> int:
    return (n + 63) // 64 * 64


def _layout(rows: int, D: int, dtype16: int):
    off_ids = HEADER_BYTES
    off_f32 = off_ids + _pad64(rows * 8)
    off_f16 = off_f32 + _pad64(rows * D * 4)
    end = off_f16 + (_pad64(rows * D * 2) if dtype16 else 0)
    return off_ids, off_f32, off_f16, end


def is_shard(path: str) -> bool:
    try:
        with open(path, "rb") as f:
            return f.read(8) == MAGIC
    except OSError:
        return False


def read_header(path: str) -> dict:
    with open(path, "rb") as f:
        raw = f.read(HEADER_BYTES)
    if len(raw) < HEADER_BYTES:
        raise ValueError(f"{path}: too short for a feature shard")
    magic, version, D, rows, dtype16, normalized = _HDR.unpack(raw)
    if magic != MAGIC:
        raise ValueError(f"{path}: not a feature shard (bad magic)")
    if version != VERSION:
        raise ValueError(f"{path}: unsupported shard version {version}")
    if dtype16 not in (DT16_NONE, DT16_F16, DT16_BF16):
        raise ValueError(f"{path}: unknown 16-bit dtype code {dtype16}")
    size = os.path.getsize(path)
    if size < _layout(rows, D, dtype16)[3]:
        raise ValueError(f"{path}: truncated shard ({size} bytes for {rows} x {D})")
    return {"rows": int(rows), "D": int(D), "dtype16": int(dtype16), "normalized": bool(normalized)}


def read_shard(path: str, mmap: bool = True):
    """-> (ids int64 [rows], feat32 float32 [rows, D], feat16 uint16 [rows, D] | None, header).
    feat16 holds the raw 16-bit patterns (view it as torch.float16 / torch.bfloat16 per header)."""
    h = read_header(path)
    rows, D = h["rows"], h["D"]
    off_ids, off_f32, off_f16, _ = _layout(rows, D, h["dtype16"])
    if mmap and rows > 0:
        ids = np.memmap(path, dtype=np.int64, mode="r", offset=off_ids, shape=(rows,))
        f32 = np.memmap(path, dtype=np.float32, mode="r", offset=off_f32, shape=(rows, D))
        f16 = (np.memmap(path, dtype=np.uint16, mode="r", offset=off_f16, shape=(rows, D))
               if h["dtype16"] else None)
    else:
        with open(path, "rb") as f:
            f.seek(off_ids)
            ids = np.frombuffer(f.read(rows * 8), dtype=np.int64).copy()
            f.seek(off_f32)
            f32 = np.frombuffer(f.read(rows * D * 4), dtype=np.float32).reshape(rows, D).copy()
            f16 = None
            if h["dtype16"]:
                f.seek(off_f16)
                f16 = np.frombuffer(f.read(rows * D * 2), dtype=np.uint16).reshape(rows, D).copy()
    return ids, f32, f16, h


def write_shard(path: str, ids, feat32: np.ndarray, feat16: Optional[np.ndarray] = None,
                dtype16: int = DT16_NONE, normalized: bool = False) -> None:
    """One-shot writer.  `feat16`: uint16 bit patterns [rows, D] of the 16-bit copy (or None)."""
    ids = np.asarray(ids, dtype=np.int64)
    feat32 = np.ascontiguousarray(feat32, dtype=np.float32)
    rows = ids.shape[0]
    D = feat32.shape[1] if feat32.ndim == 2 else 0
    if feat32.shape != (rows, D):
        raise ValueError("ids and features disagree on the number of rows")
    if (feat16 is None) != (dtype16 == DT16_NONE):
        raise ValueError("feat16 and dtype16 must be given together")
    off_ids, off_f32, off_f16, end = _layout(rows, D, dtype16)
    with open(path, "wb") as f:
        f.write(_HDR.pack(MAGIC, VERSION, D, rows, dtype16, 1 if normalized else 0))
        f.seek(off_ids)
        f.write(ids.tobytes())
        f.seek(off_f32)
        f.write(feat32.tobytes())
        if feat16 is not None:
            feat16 = np.ascontiguousarray(feat16).view(np.uint16)
            if feat16.shape != (rows, D):
                raise ValueError("feat16 must be [rows, D]")
            f.seek(off_f16)
            f.write(feat16.tobytes())
        f.truncate(end)


class FeatureWriter:
    """Incremental writer for an extract_features-style loop with an unknown-in-advance row count
    bounded by `capacity` (the file is finalised to the rows actually appended).

        w = FeatureWriter(path, D=512, capacity=len(dataset), feat_dtype=torch.float16)
        for ids, x in loader:                       # x: raw tower outputs on the GPU
            w.append(ids, model(x, None))           # normalise + cast on the device (kernel 1)
        w.close()
    """

    def __init__(self, path: str, D: int, capacity: int, feat_dtype=None, normalize: bool = True):
        import torch
        self.path, self.D, self.capacity, self.normalize = path, int(D), int(capacity), bool(normalize)
        self.torch_dtype16 = feat_dtype
        self.dtype16 = {None: DT16_NONE, torch.float16: DT16_F16, torch.bfloat16: DT16_BF16}[feat_dtype]
        self.rows = 0
        self._offs = _layout(self.capacity, self.D, self.dtype16)
        self._f = open(path, "wb+")
        self._f.truncate(self._offs[3])

    def append(self, ids, features) -> None:
        """`features`: [b, D] CUDA tensor (any float dtype).  Normalised (if requested) and cast by
        kernel (1); the fp32 and 16-bit rows are copied back and written at their final offsets."""
        import torch
        from .. import kernels as K
        b = features.shape[0]
        if features.shape[1] != self.D:
            raise ValueError(f"expected D={self.D}, got {features.shape[1]}")
        if self.rows + b > self.capacity:
            raise ValueError("FeatureWriter capacity exceeded")
        if not features.is_cuda:
            raise RuntimeError("FeatureWriter.append needs a CUDA tensor (kernel (1) has no CPU path)")
        y16, y32, _ = K.l2norm_cast(features, self.torch_dtype16, normalize=self.normalize, want_fp32=True)
        ids = np.asarray(ids.cpu() if torch.is_tensor(ids) else ids, dtype=np.int64)
        if ids.shape != (b,):
            raise ValueError("ids must be one per feature row")
        off_ids, off_f32, off_f16, _ = self._offs
        self._f.seek(off_ids + self.rows * 8)
        self._f.write(ids.tobytes())
        self._f.seek(off_f32 + self.rows * self.D * 4)
        self._f.write(y32.cpu().numpy().tobytes())
        if y16 is not None:
            self._f.seek(off_f16 + self.rows * self.D * 2)
            self._f.write(y16.view(torch.int16).cpu().numpy().tobytes())
        self.rows += b

    def close(self) -> None:
        """Compact to the rows actually written (sections move down if capacity was not reached)."""
        if self._f is None:
            return
        f, rows, D = self._f, self.rows, self.D
        if rows != self.capacity:
            _, src32, src16, _ = self._offs
            off_ids, off_f32, off_f16, end = _layout(rows, D, self.dtype16)
            for src, dst, nbytes in ((src32, off_f32, rows * D * 4),
                                     (src16, off_f16, rows * D * 2 if self.dtype16 else 0)):
                done = 0
                while done < nbytes:  # dst <= src: a forward chunked copy never overwrites unread data
                    n = min(64 << 20, nbytes - done)
                    f.seek(src + done)
                    buf = f.read(n)
                    f.seek(dst + done)
                    f.write(buf)
                    done += n
            f.truncate(end)
        f.seek(0)
        f.write(_HDR.pack(MAGIC, VERSION, D, rows, self.dtype16, 1 if self.normalize else 0))
        f.close()
        self._f = None

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()


def _load_jsonl_python(path: str, id_key: str):
    """The reference's own parse (make_topk_predictions.py:57-65): json.loads per line."""
    ids, feats = [], []
    with open(path, "r") as fin:
        for line in fin:
            line = line.strip()
            if not line:
                continue
            obj = json.loads(line)
            ids.append(obj[id_key])
            feats.append(obj["feature"])
    arr = np.array(feats, dtype=np.float32)
    if arr.ndim != 2:
        arr = arr.reshape(len(ids), -1)
    return ids, arr


def _load_jsonl_native(path: str, id_key: str, n_threads: int = 0, shard: tuple[int, int] | None = None):
    """The same parse in C++ (csrc/jsonl.cu: correctly rounded std::from_chars, lines in parallel):
    bit-identical arrays, ~10x faster per core.  Raises on anything but {"<id_key>": int, "feature": [...]}.
    `shard` = (rank, world): ids of every line, feature rows only of this rank's contiguous share
    [rows * rank // world, rows * (rank + 1) // world)."""
    import ctypes
    from .. import _lib
    lib = _lib.load()
    size = os.path.getsize(path)
    if size == 0:
        return [], np.zeros((0, 0), dtype=np.float32)
    buf = np.memmap(path, dtype=np.uint8, mode="r")
    rows, D = ctypes.c_int64(0), ctypes.c_int64(0)
    key = id_key.encode("utf-8")
    _lib.check(lib.nans_jsonl_scan(buf.ctypes.data, size, key, ctypes.byref(rows), ctypes.byref(D)))
    n = rows.value
    lo, hi = (0, n) if shard is None else (n * shard[0] // shard[1], n * (shard[0] + 1) // shard[1])
    ids = np.empty((n,), dtype=np.int64)
    feats = np.empty((hi - lo, D.value), dtype=np.float32)
    _lib.check(lib.nans_jsonl_parse_rows(buf.ctypes.data, size, key, n, D.value, lo, hi - lo, ids.ctypes.data,
                                         feats.ctypes.data, int(n_threads)))
    return ids.tolist(), feats


def load_jsonl_features(path: str, id_key: str, shard: tuple[int, int] | None = None):
    """{"<id_key>": int, "feature": [floats]} per line -> (ids list, float32 [n, D] array); the
    parsing of make_topk_predictions.py:57-65.  Native parser first; anything it does not accept
    (non-integer ids, ragged rows, other JSON) goes through json.loads exactly like the reference.
    With `shard` = (rank, world) the array holds only this rank's contiguous share of the rows (the ids
    still cover the whole file)."""
    if os.environ.get("NANS_JSONL_PYTHON", "0") != "1":
        try:
            return _load_jsonl_native(path, id_key, shard=shard)
        except Exception:
            pass
    ids, arr = _load_jsonl_python(path, id_key)
    if shard is not None:
        n = len(ids)
        arr = arr[n * shard[0] // shard[1]: n * (shard[0] + 1) // shard[1]]
    return ids, arr


def load_features(path: str, id_key: str, shard: tuple[int, int] | None = None):
    """JSONL or binary shard -> (ids, feat32 [n, D], feat16 bit patterns | None, dtype16 code).
    `shard` = (rank, world): feat32 / feat16 hold only the rows [n * rank // world, n * (rank + 1) // world)
    (a binary shard is memory-mapped, so only that slice is ever read; a JSONL file has only that
    slice's feature lists parsed)."""
    if is_shard(path):
        ids, f32, f16, h = read_shard(path)
        if shard is not None:
            n = len(ids)
            lo, hi = n * shard[0] // shard[1], n * (shard[0] + 1) // shard[1]
            f32 = f32[lo:hi]
            f16 = f16[lo:hi] if f16 is not None else None
        return ids, f32, f16, h["dtype16"]
    ids, f32 = load_jsonl_features(path, id_key, shard)
    return ids, f32, None, DT16_NONE


def jsonl_to_shard(jsonl_path: str, id_key: str, out_path: str) -> int:
    """Convert a reference-format JSONL feature file to a shard (fp32 only; host-side, no GPU)."""
    ids, f32 = load_jsonl_features(jsonl_path, id_key)
    write_shard(out_path, ids, f32)
    return len(ids)


def main(argv=None):
    import argparse
    p = argparse.ArgumentParser(description="Convert a JSONL feature file to a binary feature shard.")
    p.add_argument("--input", required=True)
    p.add_argument("--id-key", required=True, choices=["image_id", "text_id"])
    p.add_argument("--output", required=True)
    a = p.parse_args(argv)
    n = jsonl_to_shard(a.input, a.id_key, a.output)
    print(f"{n} features written to {a.output}")


if __name__ == "__main__":
    main()
