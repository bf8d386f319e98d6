#!/usr/bin/env python3
"""build_ref.py -- TEST INFRASTRUCTURE.  Builds oracle/_ref/libsosref.so from the reference's own Fortran sources
where they lie under /root/reference (never copied into the repository): oracle/f77_to_c.py translates the hot-path
source files to C, gcc compiles them with the oracle's flags (-O2 -ffp-contract=off -fno-fast-math: IEEE double/float
arithmetic, no FMA contraction -- what `gfortran -O` emits on x86-64).  The GPU box has no /root/reference; it uses the
prebuilt library that travels with the snapshot (oracle/_ref/ is git-ignored, not gpurun-ignored).

usage: python oracle/build_ref.py [--force]      exit code 0 also when the reference tree is absent (nothing to do)
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("SOS_REFERENCE_ROOT", "/root/reference")
OUT = os.path.join(HERE, "_ref")
# hot path + the input generator's Gauss angles + the "next" rows N1 / N2 of section 8(f), whose
# reference routines are then available to future parity tests (and make SOS_TRPHI's Roujean / BPDF branches callable)
FILES = ["SOS_OS.F", "SOS.F", "SOS_AGGREGATE.F", "SOS_TRPHI.F", "SOS_GLITTER.F", "SOS_SURFACE.F", "SOS_ANGLES.F",
         "SOS_ROUJEAN.F", "SOS_SURFACE_BPDF.F", "SOS_PROFIL.F", "SOS_ABSPROFILE.F",
         # N1: COEFF_ABS_CKD and the linear / spline interpolators it calls (the other routines of these two files do not
         # translate -- DATA tables, list-directed string I/O -- and are not on the path)
         "SOS_SUB_TRS.F", "SOS_AEROSOLS.F",
         # N3: Mie theory (SOS_MIE, SOS_FPHASE_MIE); SOS_GRANU and SOS_DECOMPO_LEGENDRE come from SOS_AEROSOLS.F above
         "SOS_MIE.F",
         # N1 rest: the gas atmosphere of a run (standard atmospheres are DATA tables of SOS_SUB_TRS.F)
         "SOS_PREPA_ABSPROFILE.F",
         # the whole run: the entry binding/run_sos.py calls (SOS_PROC) with SOS_PREPA_OS and the surface file names, so that the keyword
         # front end can be compared with the reference's own driver from its arguments to its result files
         "SOS_PROC.F", "SOS_PREPA_OS.F", "SOS_NOM_FIC_SURFACE.F"]


def build(force=False, verbose=True):
    lib = os.path.join(OUT, "libsosref.so")
    src_dir, inc = os.path.join(REF, "src"), os.path.join(REF, "inc", "SOS.h")
    if not os.path.isdir(src_dir):
        return lib if os.path.exists(lib) else None
    import importlib.util
    spec = importlib.util.spec_from_file_location("sos_f77_to_c", os.path.join(HERE, "f77_to_c.py"))
    t = importlib.util.module_from_spec(spec)                   # by path: sys.path stays untouched
    spec.loader.exec_module(t)
    srcs = [os.path.join(src_dir, f) for f in FILES]
    newest = max(os.path.getmtime(p) for p in srcs + [inc, os.path.join(HERE, "f77_to_c.py"), os.path.abspath(__file__)])
    if not force and os.path.exists(lib) and os.path.getmtime(lib) >= newest:
        return lib
    os.makedirs(OUT, exist_ok=True)
    defines = t.read_defines(inc)
    protos, bodies, report = [], [], []
    known = set()
    for p in srcs:
        known |= t.subroutine_names(p)
    skipped = set()
    for p in srcs:                                              # first pass: which routines translate at all
        _, _, rep = t.translate_file(p, defines, known=known)
        skipped |= {name for name, st in rep if st != "ok"}      # calls of skipped routines become run-time aborts
    known -= skipped
    for p in srcs:
        c, pr, rep = t.translate_file(p, defines, known=known, skipped=skipped)
        protos += pr
        bodies.append(c)
        report += [(os.path.basename(p),) + r for r in rep]
    csrc = os.path.join(OUT, "sosref.c")
    with open(csrc, "w") as f:
        f.write(t.PRELUDE + "\n".join(protos) + "\n\n" + "\n".join(bodies))
        # test-only access to the Fortran unit table, so that a test can hand an open unit to SOS_OUTPUT_HEADER(_POLAR_DIAG)
        f.write("\nint sosref_open_unit(int u, const char *path) { return f77_open(u, path, strlen(path), 4); }\n"
                "void sosref_close_unit(int u) { f77_close(u, 0); }\n")
    with open(os.path.join(OUT, "translation_report.txt"), "w") as f:
        for fn, name, st in report:
            f.write("%-16s %-32s %s\n" % (fn, name, st))
    cmd = ["gcc", "-O2", "-ffp-contract=off", "-fno-fast-math", "-w", "-shared", "-fPIC", "-o", lib, csrc, "-lm"]
    subprocess.run(cmd, check=True)
    os.remove(csrc)                                             # the translated sources are not kept
    if verbose:
        ok = sum(1 for r in report if r[2] == "ok")
        print("oracle/_ref: %d of %d reference subroutines translated and compiled -> %s" % (ok, len(report), lib))
    return lib


if __name__ == "__main__":
    build(force="--force" in sys.argv)
