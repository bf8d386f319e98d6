"""Codecs for the file formats on the SOS-ABS hot path.

Every channel between the reference's stages is lossy (text with 8 or 15 significant
digits, REAL*4 unformatted records); the solver must consume the *rounded* values, so the
synthetic generators push their numbers through these codecs before handing them to either
the CUDA path or the oracle.

  profile   TMP/PROFIL_TMP   SOS_PROFIL.F:1081-1085 / SOS.F:511-516   fmt 70: 2X,I5,F10.5,3(E15.8)
  angles    SOS_UsedAngles   SOS_ANGLES.F:494-506, fmt 600: I4,1X,2D21.14,1X,I4
  aerosols  Aerosols.txt     SOS_PREPA_OS.F:666-694, rows E15.8,3(1X,E15.8)
  result    SOS_Result.bin   SOS_OS.F:1571-1575   unformatted sequential, Q,U,I(-N:N) FP64
  surface   GLITTER-...      SOS_SURFACE.F:2404-2412 / SOS_OS.F:916-925  unformatted, 9 NxN REAL*4
"""
import struct

import numpy as np


def fortran_e(x, w, d, expchar="E"):
    """gfortran Ew.d / Dw.d edit descriptor: 0.dddddE+ee, correctly rounded, right-justified in w."""
    if x == 0.0:
        s = "0." + "0" * d + expchar + "+00"
    else:
        m = "%.*e" % (d - 1, abs(x))          # d.ddde+xx  (correctly rounded, ties-to-even on the binary value)
        mant, ex = m.split("e")
        digits = mant.replace(".", "")
        e = int(ex) + 1
        s = "0." + digits + expchar + ("+" if e >= 0 else "-") + "%02d" % abs(e)
        if x < 0:
            s = "-" + s
    if len(s) > w and s.startswith("0."):
        s = s[1:]                              # gfortran drops the optional leading zero when tight
    elif len(s) > w and s.startswith("-0."):
        s = "-" + s[2:]
    return s.rjust(w) if len(s) <= w else "*" * w


def fortran_f(x, w, d):
    s = "%.*f" % (d, x)
    return s.rjust(w) if len(s) <= w else "*" * w


def round_e(x, d):
    """Value of x after a write/read round trip through Ew.d (d significant digits)."""
    a = np.asarray(x, dtype=np.float64)
    flat = np.array([float("%.*e" % (d - 1, v)) for v in a.ravel()])
    out = flat.reshape(a.shape)
    return out if out.shape else float(out)


def round_f(x, d):
    a = np.asarray(x, dtype=np.float64)
    flat = np.array([float("%.*f" % (d, v)) for v in a.ravel()])
    out = flat.reshape(a.shape)
    return out if out.shape else float(out)


# ---------------------------------------------------------------- profile file
def write_profile(path, zprof, h, pcaer, pcmol):
    with open(path, "w") as f:
        for i in range(len(h)):
            f.write("  %5d%s%s%s%s\n" % (i, fortran_f(zprof[i], 10, 5), fortran_e(h[i], 15, 8),
                                        fortran_e(pcaer[i], 15, 8), fortran_e(pcmol[i], 15, 8)))


def read_profile(path):
    z, h, xa, ym = [], [], [], []
    with open(path) as f:
        for line in f:
            if not line.strip():
                continue
            z.append(float(line[7:17]))
            h.append(float(line[17:32]))
            xa.append(float(line[32:47]))
            ym.append(float(line[47:62]))
    return (np.array(z), np.array(h), np.array(xa), np.array(ym))


# ---------------------------------------------------------------- angles file
def write_angles(path, rmu_pos, ga_pos, user_flag, nb_gauss, tetas, imus, os_nb, os_ns, os_nm,
                 userfile="NO_USER_ANGLES"):
    n = len(rmu_pos)
    with open(path, "w") as f:
        f.write("NB_TOTAL_ANGLES :%4d\n" % n)
        f.write("NB_GAUSS_ANGLES :%4d\n" % nb_gauss)
        f.write("ANGLES_USERFILE :%s\n" % userfile)
        f.write("SOLAR ZENITH ANGLE :%s\n" % fortran_f(tetas, 7, 3))
        f.write("INTERNAL_IMUS :%4d\n" % imus)
        f.write("INTERNAL_OS_NB :%4d\n" % os_nb)
        f.write("INTERNAL_OS_NS :%4d\n" % os_ns)
        f.write("INTERNAL_OS_NM :%4d\n" % os_nm)
        f.write("INDEX   COS_ANGLE            WEIGHT             USER_ANGLE\n")
        for j in range(n):
            f.write("%4d %s%s %4d\n" % (j + 1, fortran_e(rmu_pos[j], 21, 14, "D"),
                                       fortran_e(ga_pos[j], 21, 14, "D"), user_flag[j]))


def read_angles(path):
    with open(path) as f:
        lines = f.read().splitlines()
    hdr = {}
    for ln in lines[:8]:
        k, v = ln.split(":", 1)
        hdr[k.strip()] = v.strip()
    n = int(hdr["NB_TOTAL_ANGLES"])
    mu, w, flag = [], [], []
    for ln in lines[9:9 + n]:
        p = ln.replace("D", "E").split()
        mu.append(float(p[1])); w.append(float(p[2])); flag.append(int(p[3]))
    return dict(nbmu=n, nb_gauss=int(hdr["NB_GAUSS_ANGLES"]), tetas=float(hdr["SOLAR ZENITH ANGLE"]),
                n0=int(hdr["INTERNAL_IMUS"]), os_nb=int(hdr["INTERNAL_OS_NB"]),
                os_ns=int(hdr["INTERNAL_OS_NS"]), os_nm=int(hdr["INTERNAL_OS_NM"]),
                rmu=np.array(mu), ga=np.array(w), user=np.array(flag))


# ---------------------------------------------------------------- aerosols file
def write_aerosols(path, a_trunc, piztr, alpha, beta, gamma, zeta, sig_ext=1.0, sig_sca=1.0, g=0.0):
    with open(path, "w") as f:
        f.write("EXTINCTION CROSS SECTION (mic^2) : %s\n" % fortran_e(sig_ext, 15, 8))
        f.write("SCATTERING CROSS SECTION (mic^2) : %s\n" % fortran_e(sig_sca, 15, 8))
        f.write("ASYMMETRY FACTOR (no truncation) : %s\n" % fortran_f(g, 9, 5))
        f.write("TRUNCATION COEFFICIENT : %s\n" % fortran_f(a_trunc, 9, 5))
        f.write("SINGLE SCATTERING ALBEDO (truncation): %s\n" % fortran_f(piztr, 9, 5))
        f.write("\nPHASE MATRIX COEFFICIENTS\n     ALPHA(K)        BETA(K)         GAMMA(K)        ZETA(K)\n")
        for k in range(len(beta)):
            f.write("%s %s %s %s\n" % tuple(fortran_e(v[k], 15, 8) for v in (alpha, beta, gamma, zeta)))


def read_aerosols(path, os_nb):
    with open(path) as f:
        lines = f.read().splitlines()
    a = float(lines[3].split(":", 1)[1])
    piztr = float(lines[4].split(":", 1)[1])
    rows = [[float(x) for x in ln.split()] for ln in lines[8:8 + os_nb + 1]]
    arr = np.array(rows)
    piz = piztr / (1 + 0.5 * a * (piztr - 1))          # SOS_PREPA_OS.F:700
    return dict(a=a, piztr=piztr, piz=piz, alpha=arr[:, 0], beta=arr[:, 1], gamma=arr[:, 2], zeta=arr[:, 3])


# ---------------------------------------------------------------- gfortran unformatted sequential
def _write_record(f, payload):
    n = len(payload)
    f.write(struct.pack("<i", n)); f.write(payload); f.write(struct.pack("<i", n))


def _read_records(path):
    recs = []
    with open(path, "rb") as f:
        while True:
            head = f.read(4)
            if len(head) < 4:
                break
            (n,) = struct.unpack("<i", head)
            recs.append(f.read(n))
            f.read(4)
    return recs


def write_result_bin(path, rec):
    """rec: [nrec, 3, 2N+1] in file order Q,U,I (SOS_OS.F:1572-1574)."""
    with open(path, "wb") as f:
        for r in np.asarray(rec, dtype="<f8"):
            _write_record(f, r.tobytes())


def read_result_bin(path, nbmu):
    w = 2 * nbmu + 1
    recs = [np.frombuffer(b, dtype="<f8").reshape(3, w) for b in _read_records(path)]
    return np.array(recs)


def write_surface_bin(path, surf):
    """surf: [nrec, 9, N(J), N(I)] REAL*4 such that surf[s, m, J-1, I-1] = R_m(I,J) (SOS_SURFACE.F:2404-2412)."""
    with open(path, "wb") as f:
        for r in np.asarray(surf, dtype="<f4"):
            _write_record(f, r.tobytes())


def read_surface_bin(path, nbmu):
    recs = [np.frombuffer(b, dtype="<f4").reshape(9, nbmu, nbmu) for b in _read_records(path)]
    return np.array(recs)
