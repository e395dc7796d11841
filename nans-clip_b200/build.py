"""Build the sm_100a shared library in-tree: nans-clip_b200/lib/libnans_clip.so.

One `nvcc` invocation per translation unit (so a touched kernel recompiles alone), then a link.
Objects are rebuilt when their source or any header is newer.  No torch headers are involved: the
library is a plain C-ABI (`include/nans_clip.h`) loaded with ctypes.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

PKG = Path(__file__).resolve().parent
CSRC = PKG / "csrc"
INCLUDE = PKG.parent / "include"
LIBDIR = PKG / "lib"
OBJDIR = PKG / "lib" / "obj"
LIB = LIBDIR / "libnans_clip.so"

SOURCES = ["api.cu", "exchange.cu", "jsonl.cu", "l2norm.cu", "smooth.cu", "strip_fwd.cu", "strip_bwd.cu", "topk.cu"]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-std=c++17", "-lineinfo",
    "-Xcompiler", "-fPIC",
    "--expt-relaxed-constexpr",
    "-Xptxas", "-v",
] + os.environ.get("NANS_EXTRA_NVCC_FLAGS", "").split()   # bring-up experiments only (e.g. -DNANS_NP_CF_SHFL)


def _nvcc() -> str:
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        raise RuntimeError("nvcc not found: the CUDA extension cannot be built")
    return nvcc


def _stale(target: Path, deps: list[Path]) -> bool:
    if not target.exists():
        return True
    t = target.stat().st_mtime
    return any(d.stat().st_mtime > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> Path:
    """Serialised across processes (every rank of a torchrun job may get here at once on a fresh
    checkout): an exclusive flock on lib/.build.lock is held for the whole build, objects and the
    library are written under temporary names and renamed into place, so a concurrent reader never
    maps a half-written file."""
    import fcntl

    OBJDIR.mkdir(parents=True, exist_ok=True)
    with open(LIBDIR / ".build.lock", "w") as lock:
        fcntl.flock(lock, fcntl.LOCK_EX)
        try:
            return _build_locked(force, verbose)
        finally:
            fcntl.flock(lock, fcntl.LOCK_UN)


def _build_locked(force: bool, verbose: bool) -> Path:
    nvcc = _nvcc()
    headers = sorted(CSRC.glob("*.cuh")) + sorted(INCLUDE.glob("*.h")) + [Path(__file__)]
    jobs = []
    objs = []
    for name in SOURCES:
        src = CSRC / name
        obj = OBJDIR / (src.stem + ".o")
        objs.append(obj)
        if force or _stale(obj, [src] + headers):
            jobs.append((src, obj))

    def compile_one(job):
        src, obj = job
        tmp = obj.with_suffix(f".{os.getpid()}.tmp.o")
        cmd = [nvcc, *NVCC_FLAGS, "-I", str(INCLUDE), "-c", str(src), "-o", str(tmp)]
        r = subprocess.run(cmd, capture_output=True, text=True)
        log = OBJDIR / (src.stem + ".ptxas.log")   # untracked (registers / spills of the last build)
        log.write_text(r.stdout + r.stderr)
        if r.returncode != 0:
            tmp.unlink(missing_ok=True)
            raise RuntimeError(f"nvcc failed for {src.name}:\n{r.stdout}\n{r.stderr}")
        os.replace(tmp, obj)
        if verbose:
            print(r.stderr)
        return obj

    if jobs:
        with ThreadPoolExecutor(max_workers=min(len(jobs), os.cpu_count() or 4)) as ex:
            list(ex.map(compile_one, jobs))
    if jobs or force or _stale(LIB, objs):
        tmp = LIB.with_name(f".{LIB.name}.{os.getpid()}.tmp")
        cmd = [nvcc, "-shared", "-o", str(tmp), *map(str, objs), "-cudart", "static"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            tmp.unlink(missing_ok=True)
            raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
        os.replace(tmp, LIB)
    return LIB


if __name__ == "__main__":
    path = build(force="--force" in sys.argv, verbose="-v" in sys.argv)
    print(path)
