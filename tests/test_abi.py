"""CPU checks of the drop-in boundary: libsosgpu.so loads, exports every symbol include/sosgpu.h declares,
and refuses to compute without a GPU (no fallback)."""
import ctypes as C
import importlib
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    src = open(os.path.join(ROOT, "include", "sosgpu.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    names = re.findall(r"\b(sosgpu_[a-z_0-9]+|sos_os_|sos_aggregate_|sos_|sos_glitter_|sos_trphi_|sos_trphi_option_)\s*\(", src)
    return sorted(set(names))


def test_header_symbols_exported():
    api = importlib.import_module("radiativetransfer-sos_b200.api")
    lib = api.load_library()
    syms = _declared_symbols()
    assert len(syms) >= 15
    for s in syms:
        assert hasattr(lib, s), "libsosgpu.so does not export %s declared in include/sosgpu.h" % s


def test_no_cpu_fallback():
    """Without a CUDA device the context cannot be created and the Python host layer raises."""
    api = importlib.import_module("radiativetransfer-sos_b200.api")
    lib = api.load_library()
    lib.sosgpu_device_count.restype = C.c_int
    if lib.sosgpu_device_count() > 0:
        pytest.skip("a GPU is present")
    h = C.c_void_p()
    assert lib.sosgpu_create(C.byref(h), 0) == api.SOSGPU_ERR_NO_DEVICE
    with pytest.raises(RuntimeError):
        api.Solver(0)
    # compute entries reject a null context instead of computing on the host
    assert lib.sosgpu_batch_run(None, None, 1, 1, 0, None, None) == api.SOSGPU_ERR_NO_DEVICE


def test_product_does_not_import_oracle():
    """The product package must never reach into oracle/ (test infrastructure)."""
    pdir = os.path.join(ROOT, "radiativetransfer-sos_b200")
    for dp, _, files in os.walk(pdir):
        for f in files:
            if f.endswith((".py", ".cu", ".h", ".cpp")):
                txt = open(os.path.join(dp, f), errors="ignore").read()
                assert "liboracle" not in txt and "sos_oracle" not in txt and "from oracle" not in txt, f
