"""Build libvfi.so (the C-ABI CUDA library) in-tree with nvcc for sm_100a.

nvcc cross-compiles without a GPU, so this runs in the CPU container; the built .so travels to the
GPU box with the repo snapshot (it is git-ignored, not gpurun-ignored).
"""
from __future__ import annotations

import hashlib
import os
import shutil
import subprocess
from pathlib import Path

PKG = Path(__file__).resolve().parent
CSRC = PKG / "csrc"
LIB_DIR = PKG / "_lib"
LIB_PATH = LIB_DIR / "libvfi.so"
STAMP = LIB_DIR / "libvfi.stamp"

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC",
]
OBJ_DIR = LIB_DIR / "obj"


def _units() -> list[Path]:
    """The translation units of libvfi.so (compiled in parallel, then linked)."""
    return sorted(CSRC.glob("api_*.cu"))


def _sources() -> list[Path]:
    return (sorted(CSRC.glob("*.cu")) + sorted(CSRC.glob("*.cuh")) + sorted(CSRC.glob("*.h"))
            + [PKG.parent / "include" / "vfi.h"])


def _digest() -> str:
    h = hashlib.sha256()
    for p in _sources():
        h.update(p.name.encode())
        h.update(p.read_bytes())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def find_nvcc() -> str | None:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and Path(cand).exists():
            return cand
    return None


def build_lib(force: bool = False, verbose: bool = False) -> Path:
    """Compile csrc/api_*.cu -> _lib/obj/*.o (one nvcc per unit, in parallel) and link _lib/libvfi.so, unless an
    up-to-date build exists."""
    LIB_DIR.mkdir(exist_ok=True)
    digest = _digest()
    if not force and LIB_PATH.exists() and STAMP.exists() and STAMP.read_text().strip() == digest:
        return LIB_PATH
    nvcc = find_nvcc()
    if nvcc is None:
        raise RuntimeError("nvcc not found: cannot build libvfi.so (set NVCC or install the CUDA toolkit)")
    OBJ_DIR.mkdir(exist_ok=True)
    procs = []
    for src in _units():
        obj = OBJ_DIR / (src.stem + ".o")
        cmd = [nvcc, *NVCC_FLAGS, "-c", "-o", str(obj), str(src)]
        if verbose:
            cmd[1:1] = ["-Xptxas", "-v"]
        procs.append((src, obj, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)))
    objs, errors = [], []
    for src, obj, proc in procs:
        out, err = proc.communicate()
        if proc.returncode != 0:
            errors.append(f"nvcc failed on {src.name}:\n{out}{err}")
        elif verbose:
            print(f"== {src.name}\n{err}")
        objs.append(str(obj))
    if errors:
        raise RuntimeError("\n".join(errors))
    link = subprocess.run([nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", str(LIB_PATH), *objs],
                          capture_output=True, text=True)
    if link.returncode != 0:
        raise RuntimeError("nvcc link failed:\n" + link.stdout + link.stderr)
    STAMP.write_text(digest)
    return LIB_PATH


if __name__ == "__main__":
    print(build_lib(force=True, verbose=True))
