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
    names = re.findall(r"\b(sosgpu_[a-z_0-9]+|sos_[a-z_0-9]*_|sos_|read_ckd_coeff_)\s*\(", src)
    return sorted(set(names))


def test_header_symbols_exported():
    api = importlib.import_module("radiativetransfer-sos_b200.api")
    lib = api.load_library()
    syms = _declared_symbols()
    assert len(syms) >= 15
    for s in ("sos_", "sos_os_", "sos_aggregate_", "sos_glitter_", "sos_trphi_", "sos_trphi_option_", "sos_absprofile_", "sos_profile_",
              "read_ckd_coeff_", "sos_mie_", "sos_granu_", "sos_decompo_legendre_", "sosgpu_aerosols", "sosgpu_mie"):
        assert s in syms, s
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


def test_ctypes_structs_match_header(tmp_path):
    """The ctypes mirrors in api.py must have the layout a C compiler gives include/sosgpu.h (sizes and offsets)."""
    import subprocess
    api = importlib.import_module("radiativetransfer-sos_b200.api")
    pairs = {"sosgpu_optics": api.COptics, "sosgpu_term": api.CTerm, "sosgpu_term_out": api.CTermOut,
             "sosgpu_group_out": api.CGroupOut, "sosgpu_stats": api.CStats, "sosgpu_direct_models": api.CDirectModels,
             "sosgpu_ckd": api.CCkd, "sosgpu_gas_profile": api.CGasProfile, "sosgpu_profile_term": api.CProfileTerm,
             "sosgpu_aer_component": api.CAerComponent, "sosgpu_aer_model": api.CAerModel}
    lines = ['#include <stdio.h>', '#include <stddef.h>', '#include "sosgpu.h"', 'int main(void) {']
    for cname, cls in pairs.items():
        lines.append('  printf("%s size %%zu\\n", sizeof(%s));' % (cname, cname))
        for fname, _ in cls._fields_:
            lines.append('  printf("%s %s %%zu\\n", offsetof(%s, %s));' % (cname, fname, cname, fname.rstrip("_")))
    lines += ['  return 0;', '}']
    src = tmp_path / "layout.c"
    src.write_text("\n".join(lines))
    exe = tmp_path / "layout"
    subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), "-o", str(exe), str(src)], check=True)
    out = subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.split("\n")
    got = {}
    for ln in out:
        if ln.strip():
            a, b, c = ln.split()
            got[(a, b)] = int(c)
    for cname, cls in pairs.items():
        assert got[(cname, "size")] == C.sizeof(cls), cname
        for fname, _ in cls._fields_:
            assert got[(cname, fname)] == getattr(cls, fname).offset, (cname, fname)


def test_gfortran_shims_refuse_without_gpu(tmp_path):
    """The F77 drop-in symbols have no CPU path either: without a CUDA device they return IER = -1."""
    import numpy as np
    api = importlib.import_module("radiativetransfer-sos_b200.api")
    lib = api.load_library()
    lib.sosgpu_device_count.restype = C.c_int
    if lib.sosgpu_device_count() > 0:
        pytest.skip("a GPU is present")
    ip = lambda v: C.byref(C.c_int(v))
    dp = lambda v: C.byref(C.c_double(v))
    fstr = lambda s: C.create_string_buffer(s.encode().ljust(500), 500)
    P = lambda a: a.ctypes.data_as(C.POINTER(C.c_double))
    L500 = C.c_size_t(500)
    rmu, ga = np.zeros(161), np.zeros(161)
    rmu[80 + 1:80 + 5] = [0.9, 0.7, 0.5, 0.3]
    rmu[80 - 4:80] = -rmu[80 + 1:80 + 5][::-1]
    ier = C.c_int(0)
    out = str(tmp_path / "GLITTER.bin")
    lib.sos_glitter_(ip(4), P(rmu), P(ga), dp(2.0), dp(1.34), ip(8), ip(8), ip(16), fstr("a"), fstr("b"), fstr("c"),
                     fstr(out), ip(0), C.byref(ier), L500, L500, L500, L500)
    assert ier.value == -1 and not os.path.exists(out)
    h = np.linspace(0, 0.1, 601)
    zero = np.zeros(601)
    co = np.zeros(201)
    em, ep = C.c_double(0), C.c_double(0)
    ier = C.c_int(0)
    lib.sos_os_(ip(4), P(rmu), P(ga), ip(8), ip(10), fstr("none"), fstr(str(tmp_path / "OS.bin")), ip(1), dp(30.0), dp(0.0),
                ip(0), ip(0), dp(1.34), P(h), P(zero), P(zero + 1), P(zero), dp(0.0279), P(co), P(co), P(co), P(co),
                dp(-1.0), ip(10), ip(2), ip(1), ip(0), ip(6), C.byref(em), C.byref(ep), C.byref(ier), L500, L500)
    assert ier.value == -1 and not os.path.exists(str(tmp_path / "OS.bin"))
    # the profile chain: SOS_PROFILE / SOS_ABSPROFILE shims (INTEGER*2 arguments are short)
    sp = lambda v: C.byref(C.c_short(v))
    altabs, tabs = np.linspace(120.0, 0.0, 50), np.zeros(50)
    nt, ier = C.c_int(0), C.c_int(0)
    prof = str(tmp_path / "PROFIL.txt")
    lib.sos_profile_(sp(1), dp(0.1), dp(8.0), dp(0.2), dp(2.0), dp(0.0), dp(0.0), sp(7), P(altabs), P(tabs), ip(0), ip(0),
                     fstr(prof), C.byref(nt), C.byref(ier), L500)
    assert ier.value == -1 and not os.path.exists(prof)
    assert lib.sosgpu_profile_chain(None, None, None, None, 1, 1, None, None, None, None, None, None, None) == api.SOSGPU_ERR_NO_DEVICE
    assert lib.sosgpu_absprofile(None, None, None, None, 1, None, None) == api.SOSGPU_ERR_NO_DEVICE
    assert lib.sosgpu_profile(None, None, None, None, 1, 1, None, None, None, None, None, None) == api.SOSGPU_ERR_NO_DEVICE
    # aerosol optics: SOS_MIE / SOS_GRANU / SOS_DECOMPO_LEGENDRE shims and the batched entries
    mu, w = np.zeros(201), np.zeros(201)
    mu[101:105], w[101:105] = [0.3, 0.5, 0.85, 0.97], 0.25
    mu[96:100], w[96:100] = -mu[101:105][::-1], 0.25
    ier = C.c_int(0)
    fmie = str(tmp_path / "MIE.bin")
    lib.sos_mie_(ip(4), P(mu), P(w), dp(1.4), dp(-0.01), dp(0.0001), dp(5.0), fstr(fmie), fstr("NO_LOG_FILE"), C.byref(ier), L500, L500)
    assert ier.value == -1 and not os.path.exists(fmie)
    ier = C.c_int(0)
    it = C.c_int(1)
    ph = np.ones(201)
    lib.sos_decompo_legendre_(C.byref(it), ip(0), ip(4), P(mu), P(w), ip(8), P(ph), P(ph.copy()), P(ph), P(ph), P(ph), dp(0.0), dp(0.0),
                              P(co), P(co), P(co), P(co), P(co), P(co), C.byref(ier))
    assert ier.value == -1
    assert lib.sosgpu_aerosols(None, 4, None, None, 1, None, 0, None, 8, None, None, None, None, None, None, None) == api.SOSGPU_ERR_NO_DEVICE
    lib.sosgpu_mie.argtypes = [C.c_void_p, C.c_int, C.c_void_p] + [C.c_double] * 4 + [C.c_int] + [C.c_void_p] * 6
    assert lib.sosgpu_mie(None, 4, None, 1.4, -0.01, 0.0001, 5.0, 0, None, None, None, None, None, None) == api.SOSGPU_ERR_NO_DEVICE
    lib.sosgpu_mie_count.argtypes = [C.c_double, C.c_double]
    assert lib.sosgpu_mie_count(0.0001, 200.0) == 4000 and lib.sosgpu_mie_count(0.0001, 5000.0) == -1     # host-side grid arithmetic only
