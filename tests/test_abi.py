"""CPU-side checks of the drop-in boundary: the C-ABI library builds for sm_100a without a GPU, loads,
and exports every symbol include/snvc_b200.h declares (no compute calls here); argument errors are
reported through the int status + snvc_last_error() contract; the product never imports the oracle."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "snvc_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(snvc_[a-z0-9_]+)\s*\(", src)))


@pytest.fixture(scope="module")
def lib():
    from snvc_b200 import build, _lib
    build.build()
    return _lib.lib()


def test_library_exports_every_declared_symbol(lib):
    from snvc_b200 import _lib
    names = _declared()
    assert len(names) >= 18
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/snvc_b200.h but not exported"
    assert sorted(_lib.EXPORTS) == names, "ctypes binding table and header disagree"
    assert lib.snvc_version() == 1


def test_argument_errors_use_status_and_last_error(lib):
    # no GPU needed: argument validation happens before any CUDA call
    st = lib.snvc_cost_volume_fwd(None, None, None, None, 1, 32, 8, 8, 4, 3, 0, 0, 0, None)   # 8 % 3 != 0
    assert st < 0 and b"multiples of downsample" in lib.snvc_last_error()
    st = lib.snvc_cost_volume_fwd(None, None, None, None, 0, 32, 8, 8, 4, 1, 0, 0, 0, None)   # empty output: early return
    assert st == 0
    from snvc_b200._lib import ConvDesc
    d = ConvDesc(N=1, Cin=24, Cout=32, Di=4, Hi=4, Wi=4, Do=4, Ho=4, Wo=4, kernel=3, stride=1, pad=1, dilation=1)
    st = lib.snvc_conv3d_fwd(ctypes.c_void_p(16), ctypes.c_void_p(16), None, None, None, ctypes.c_void_p(16),
                             ctypes.byref(d), None)
    assert st < 0 and b"Cin must be" in lib.snvc_last_error()
    assert lib.snvc_conv3d_packed_weight_bytes(32, 1, 3) == 27 * 16 * 32 * 2


def test_cpu_tensors_are_rejected_like_the_reference():
    import torch
    from snvc_b200.extension.build_cost_volume import build_cost_volume
    x = torch.zeros(1, 8, 4, 8)
    with pytest.raises(RuntimeError, match="Not implemented on the CPU"):     # BuildCostVolume.cpp:26
        build_cost_volume(x, x, torch.zeros(1, 2), 1)


def test_product_does_not_import_the_oracle():
    pat = re.compile(r"^\s*(from|import)\s+oracle\b", re.M)
    for dirpath, _, files in os.walk(os.path.join(ROOT, "snvc_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                assert not pat.search(open(os.path.join(dirpath, f)).read()), f"{f} imports oracle/"
