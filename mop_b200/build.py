"""In-tree build of libmop_b200.so with nvcc for sm_100a (cross-compiles without a GPU).

One translation unit per kernel family (``csrc/abi_*.cu``), compiled in parallel to
``csrc/_obj/*.o`` and linked into ``mop_b200/libmop_b200.so``; an object is rebuilt
only when its source, a header or the build flags are newer.
"""
from __future__ import annotations

import concurrent.futures as cf
import glob
import hashlib
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(CSRC, "_obj")
OUT = os.environ.get("MOP_B200_LIB") or os.path.join(HERE, "libmop_b200.so")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC", "-diag-suppress", "177",
]


def _headers():
    return glob.glob(os.path.join(CSRC, "*.cuh")) + glob.glob(os.path.join(CSRC, "*.h")) + \
        [os.path.join(HERE, "..", "include", "mop_b200.h")]


def _units():
    return sorted(glob.glob(os.path.join(CSRC, "abi_*.cu")))


def _flag_tag(extra) -> str:
    return hashlib.sha1(" ".join(NVCC_FLAGS + list(extra)).encode()).hexdigest()[:10]


def _obj_for(src: str, tag: str) -> str:
    return os.path.join(OBJ, os.path.basename(src)[:-3] + f".{tag}.o")


def _stale_obj(src: str, obj: str) -> bool:
    if not os.path.exists(obj):
        return True
    t = os.path.getmtime(obj)
    return any(os.path.getmtime(s) > t for s in [src] + _headers())


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile csrc/abi_*.cu into mop_b200/libmop_b200.so.  Returns the path."""
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    extra = os.environ.get("MOP_NVCC_EXTRA", "").split()   # e.g. -DMOP_PHASE_TIMING (development only)
    tag = _flag_tag(extra)
    os.makedirs(OBJ, exist_ok=True)
    units = _units()
    todo = [s for s in units if force or _stale_obj(s, _obj_for(s, tag))]
    objs = [_obj_for(s, tag) for s in units]
    if not todo and os.path.exists(OUT) and all(os.path.getmtime(o) <= os.path.getmtime(OUT) for o in objs):
        return OUT

    def compile_one(src):
        cmd = [nvcc, *NVCC_FLAGS, *extra, "-c", "-o", _obj_for(src, tag), src]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
        r = subprocess.run(cmd, capture_output=True, text=True)
        return src, r

    with cf.ThreadPoolExecutor(max_workers=min(len(todo), os.cpu_count() or 1) or 1) as ex:
        for src, r in ex.map(compile_one, todo):
            if r.returncode != 0:
                raise RuntimeError(f"nvcc failed on {os.path.basename(src)}:\n{r.stdout}\n{r.stderr}")
            if verbose:
                print(f"== {os.path.basename(src)}\n{r.stderr}", file=sys.stderr)
    # no link-time libcuda: the one driver entry point (cuTensorMapEncodeTiled) is resolved at run time
    r = subprocess.run([nvcc, "-shared", "-o", OUT, *objs], capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
