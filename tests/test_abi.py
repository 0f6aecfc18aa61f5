"""CPU tests of the drop-in boundary: libmcxgpu.so loads, exports every symbol that
include/mcx_gpu.h declares, and refuses to work (loudly) without a CUDA device."""
import ctypes as C
import os
import re

import pytest

from conftest import ROOT


def _declared():
    src = open(os.path.join(ROOT, "include", "mcx_gpu.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(mcx_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    import mccortex_b200 as M
    L = C.CDLL(M.lib_path())
    names = _declared()
    assert len(names) >= 15
    for n in names:
        assert hasattr(L, n), "libmcxgpu.so does not export " + n


def test_no_cpu_fallback_without_device():
    import mccortex_b200 as M
    if M.device_count() > 0:
        pytest.skip("a CUDA device is present")
    with pytest.raises(M.McxError) as e:
        M.Graph(31, 1, 1024)
    assert e.value.status == "MCX_ERR_NO_DEVICE"


def test_product_does_not_reference_oracle():
    """The product (library sources, host driver, python mirror) must never touch oracle/."""
    bad = []
    for root, _, files in os.walk(os.path.join(ROOT, "mccortex_b200")):
        for f in files:
            if f.endswith((".cu", ".cuh", ".h", ".c", ".py", "Makefile")):
                txt = open(os.path.join(root, f), errors="replace").read()
                if re.search(r"liboracle|oracle/|from oracle|import oracle|\borc_|_ref/", txt):
                    bad.append(os.path.join(root, f))
    assert not bad, bad
