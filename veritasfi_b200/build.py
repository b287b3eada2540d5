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
    "-shared", "-Xcompiler", "-fPIC",
]


def _sources() -> list[Path]:
    return sorted(CSRC.glob("*.cu")) + sorted(CSRC.glob("*.cuh")) + [PKG.parent / "include" / "vfi.h"]


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
    """Compile csrc/vfi_api.cu -> _lib/libvfi.so unless an up-to-date build exists."""
    LIB_DIR.mkdir(exist_ok=True)
    digest = _digest()
    if not force and LIB_PATH.exists() and STAMP.exists() and STAMP.read_text().strip() == digest:
        return LIB_PATH
    nvcc = find_nvcc()
    if nvcc is None:
        raise RuntimeError("nvcc not found: cannot build libvfi.so (set NVCC or install the CUDA toolkit)")
    cmd = [nvcc, *NVCC_FLAGS, "-o", str(LIB_PATH), str(CSRC / "vfi_api.cu")]
    if verbose:
        cmd.insert(1, "-Xptxas")
        cmd.insert(2, "-v")
    proc = subprocess.run(cmd, capture_output=True, text=True)
    if proc.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + proc.stdout + proc.stderr)
    if verbose:
        print(proc.stderr)
    STAMP.write_text(digest)
    return LIB_PATH


if __name__ == "__main__":
    print(build_lib(force=True, verbose=True))
