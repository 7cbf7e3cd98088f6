"""CPU tests: the C-ABI library loads and exports every symbol include/gpss.h declares; no compute calls."""
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    txt = open(os.path.join(ROOT, "include", "gpss.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(gpss_[A-Za-z0-9_]+)\s*\(", txt)))


def test_header_symbols_are_exported(gpss):
    import ctypes
    lib = ctypes.CDLL(gpss.LIB_PATH)
    declared = _declared_symbols()
    assert len(declared) >= 20
    for name in declared:
        assert hasattr(lib, name), "libgpss.so does not export %s" % name
    # and the ctypes harness binds exactly the declared set
    assert sorted(gpss.exported_symbols()) == declared


def test_no_cuda_means_loud_failure(gpss):
    """There is no CPU fallback: without a device every compute entry point reports an error."""
    import numpy as np
    try:
        ndev = gpss.device_count()
    except gpss.GpssError:
        ndev = 0
    if ndev > 0:
        pytest.skip("a CUDA device is present")
    with pytest.raises(gpss.GpssError):
        gpss.GpssModel(np.zeros((10, 3)), np.zeros(10))


def test_product_does_not_import_oracle():
    """The oracle is test infrastructure: nothing under gp_ss_ak_b200/ may reference it."""
    pkg = os.path.join(ROOT, "gp_ss_ak_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp", ".hpp")) or f == "Makefile":
                txt = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "import oracle" not in txt and "from oracle" not in txt and "oracle/" not in txt.replace("oracle/gpss_oracle.py", ""), f
