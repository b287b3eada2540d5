"""The C-ABI shared library: loads without a GPU, exports every symbol include/vfi.h declares, validates
arguments, and fails loudly (no CPU fallback) when there is no CUDA device."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from veritasfi_b200 import _native as N
from veritasfi_b200.build import LIB_PATH

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "vfi.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(vfi_[a-z0-9_]+)\s*\(", text)))


def test_library_is_built_in_tree_and_loads():
    assert LIB_PATH.exists()
    assert str(LIB_PATH).startswith(ROOT)
    lib = N.load()
    assert lib.vfi_abi_version() == 1


def test_every_declared_symbol_is_exported_and_bound():
    declared = _declared_symbols()
    assert len(declared) >= 30
    lib = C.CDLL(str(LIB_PATH))
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/vfi.h but not exported"
    assert sorted(N.EXPORTED_SYMBOLS) == declared, "ctypes prototypes out of sync with include/vfi.h"


def test_no_torch_or_cpp_types_in_the_header():
    text = open(os.path.join(ROOT, "include", "vfi.h")).read()
    assert "at::" not in text and "torch::" not in text
    assert "std::" not in text and "template" not in text
    assert 'extern "C"' in text


def test_sass_contains_blackwell_tensor_and_tma_instructions():
    import shutil
    import subprocess
    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(cuobjdump):
        pytest.skip("cuobjdump not available")
    sass = subprocess.run([cuobjdump, "-sass", str(LIB_PATH)], capture_output=True, text=True).stdout
    assert "UTCHMMA" in sass, "tcgen05.mma missing from SASS"
    assert "LDTM" in sass, "tcgen05.ld missing from SASS"
    assert "UTMALDG" in sass, "TMA loads missing from SASS"
    assert "sm_100a" in sass
    assert "HMMA.16816" not in sass, "legacy mma.sync path present"


@pytest.mark.skipif(N.device_count() > 0, reason="only meaningful without a GPU")
def test_compute_entries_fail_loudly_without_a_device():
    from veritasfi_b200 import faiss_compat, fusion
    with pytest.raises(N.NoDeviceError):
        faiss_compat.IndexFlatIP(16)
    x = np.ones((2, 4), np.float32)
    with pytest.raises(N.NoDeviceError):
        faiss_compat.normalize_L2(x)
    assert (x == 1).all()                       # untouched: nothing ran on the CPU instead
    with pytest.raises(N.NoDeviceError):
        fusion.rrf(np.zeros((1, 2, 3), np.int64), 2)
    ex = C.c_void_p()
    assert N.load().vfi_exchange_create(0, 0, 2, 8, 8, C.byref(ex)) == N.ERR_NO_DEVICE   # no CPU stand-in for peer windows
    from veritasfi_b200.bm25_compat import GpuPostings
    with pytest.raises(N.NoDeviceError):
        GpuPostings(np.array([0, 1]), np.array([0], np.int32), np.array([1.0], np.float32), 1)


def test_argument_validation_happens_before_any_device_work():
    lib = N.load()
    h = C.c_void_p()
    assert lib.vfi_index_create(0, N.STORE_BF16, 0, C.byref(h)) == N.ERR_INVALID
    assert lib.vfi_index_create(64, 7, 0, C.byref(h)) == N.ERR_INVALID
    assert b"store_dtype" in lib.vfi_last_error()
    assert lib.vfi_index_create(64, N.STORE_BF16, 0, None) == N.ERR_INVALID
    assert lib.vfi_index_search(None, None, 1, 1, None, None, 0, None) == N.ERR_INVALID
    assert lib.vfi_merge_topk(None, None, 1, 1, 1, 1, None, None, 0, 0, None) == N.ERR_INVALID
    ex = C.c_void_p()
    assert lib.vfi_exchange_create(0, 0, 17, 8, 8, C.byref(ex)) == N.ERR_INVALID      # world > 16
    assert lib.vfi_exchange_create(0, 2, 2, 8, 8, C.byref(ex)) == N.ERR_INVALID       # rank >= world
    assert lib.vfi_exchange_create(0, 0, 2, 8, N.MAX_K + 1, C.byref(ex)) == N.ERR_INVALID
    assert lib.vfi_exchange_merge(None, None, None, 1, 1, 1, None, None, None) == N.ERR_INVALID
    assert lib.vfi_exchange_destroy(None) == N.OK
    assert lib.vfi_index_ntotal(None) == 0
    assert lib.vfi_index_destroy(None) == N.OK
    assert lib.vfi_bm25_destroy(None) == N.OK
    assert lib.vfi_launch_count() >= 0


def test_faiss_facade_rejects_bad_arrays_like_faiss():
    from veritasfi_b200 import faiss_compat
    with pytest.raises(TypeError):
        faiss_compat.normalize_L2(np.ones((2, 2), np.float64))
    with pytest.raises(ValueError):
        faiss_compat.normalize_L2(np.ones((4,), np.float32))
    with pytest.raises(ValueError):
        faiss_compat.normalize_L2(np.ones((4, 4), np.float32)[:, ::2])
    faiss_compat.normalize_L2(np.ones((0, 4), np.float32))   # empty is a no-op, no device needed
