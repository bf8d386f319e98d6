"""Shared helpers of the parity tests."""
import numpy as np

RTOL = 1e-9     # north_star: Stokes I/Q/U within 1e-9 relative ...
ATOL = 1e-12    # ... with a 1e-12 absolute floor


def assert_stokes_close(got, ref, what=""):
    got, ref = np.asarray(got), np.asarray(ref)
    assert got.shape == ref.shape, (what, got.shape, ref.shape)
    err = np.abs(got - ref) - (RTOL * np.abs(ref) + ATOL)
    worst = np.unravel_index(np.argmax(err), err.shape) if err.size else None
    assert err.size == 0 or err.max() <= 0, "%s: |diff| %.3e at %s (got %.17g ref %.17g)" % (
        what, np.abs(got - ref)[worst], worst, got[worst], ref[worst])


def oracle_term(orc, o, t, want_trans=False):
    """SOS (SOS.F:340) through the oracle for one synth.Term."""
    return orc.sos(t.nt, o.zout, o.igmax, o.ipolar, o.ron, o.ind_surf, o.rho, o.imat_surf, o.ifresnel, o.surf, o.n0,
                   o.piz, o.piztr, o.a_trunc, o.rmu, o.ga, o.tetas, o.os_nb, o.nbmu, o.alpha, o.beta, o.gamma, o.zeta,
                   t.zprof, t.h, t.pcaer, t.pcmol, want_trans)


def pack_matrix(ker, ga, nbmu, rayleigh=None):
    """Matrix form of SOS_FSOURCE_ORDREIG (SOS_OS.F:2894-2905), numpy restatement of the table in DESIGN.md.
    Rows/cols ordered [d][stokes][k]; returns M with J = M @ X (the 0.5*GA(j) factor included).
    ker: dict of kernels K[k+N, j+N] = K(j,k) (oracle.noyaux convention).  rayleigh: dict(beta0,beta2,gamma2,alpha2)
    -> build the molecular part from the l=2 rows instead."""
    N = nbmu

    def K(name, a, b):
        if rayleigh is None:
            return ker[name][b + N, a + N]
        xp, xr, xt = ker["xpl"], ker["xrl"], ker["xtl"]
        r = rayleigh
        return {"bp": r["beta0"] + r["beta2"] * xp[a + N] * xp[b + N],
                "gr": r["gamma2"] * xp[a + N] * xr[b + N],
                "gt": r["gamma2"] * xp[a + N] * xt[b + N],
                "arr": r["alpha2"] * xr[a + N] * xr[b + N],
                "art": r["alpha2"] * xt[a + N] * xr[b + N],
                "att": r["alpha2"] * xt[a + N] * xt[b + N]}[name]

    M = np.zeros((6 * N, 6 * N))
    for dout in range(2):
        for so in range(3):
            for k in range(1, N + 1):
                ro = dout * 3 * N + so * N + k - 1
                for din in range(2):
                    same = dout == din
                    for si in range(3):
                        for j in range(1, N + 1):
                            co = din * 3 * N + si * N + j - 1
                            sign = 1.0
                            if so == 0:
                                if si == 0: name, a, b = "bp", j, (k if same else -k)
                                elif si == 1: name, a, b = "gr", k, (j if same else -j)
                                else: name, a, b, sign = "gt", k, (j if same else -j), (-1.0 if dout == 0 else 1.0)
                            elif so == 1:
                                if si == 0: name, a, b = "gr", j, (k if same else -k)
                                elif si == 1: name, a, b = "arr", j, (k if same else -k)
                                else: name, a, b, sign = "art", j, (k if same else -k), (-1.0 if din == 0 else 1.0)
                            else:
                                if si == 0: name, a, b, sign = "gt", j, (k if same else -k), (-1.0 if din == 0 else 1.0)
                                elif si == 1: name, a, b, sign = "art", k, (j if same else -j), (-1.0 if dout == 0 else 1.0)
                                else: name, a, b = "att", j, (k if same else -k)
                            M[ro, co] = 0.5 * ga[j + N] * sign * K(name, a, b)
    return M


def fields_to_packed(i1, q1, u1, nbmu):
    """[2N+1, L] reference fields -> [6N, L] packed rows [d][stokes][k]."""
    N = nbmu
    rows = []
    for d in range(2):
        for f in (i1, q1, u1):
            for k in range(1, N + 1):
                rows.append(f[(k if d == 0 else -k) + N])
    return np.array(rows)
