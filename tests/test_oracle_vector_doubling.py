"""Independent pin of the POLARIZED multiple-scattering field: a vector (I, Q, U) adding-doubling solver built from
first principles against the oracle's successive orders.

Nothing here comes from the reference: the Rayleigh phase matrix is formed from the dipole amplitude matrix
e_a(n) . e_b(n') in the meridian-plane bases (no rotation-angle formulas, no generalized spherical functions), Fourier
transformed numerically in azimuth, and the layer is built by doubling with all four operators (R, T, R*, T*) carried
explicitly, so no symmetry relation is assumed.  The oracle side runs SOS_OS in full (12 scattering orders, geometric tail,
stop tests) with the l = 2 polarization kernels of SOS_NOYAUX and the sign table of SOS_FSOURCE_ORDREIG.
Conventions that may legitimately differ are global signs of Q and U; U comes out with the opposite sign, I and Q equal."""
import numpy as np
import pytest


def mueller_rayleigh(ct, phi, ctp, delta):
    """3x3 (I,Q,U) phase matrix for scattering from (theta', phi'=0) into (theta, phi); Stokes referred to the meridian
    planes with axes (theta_hat, phi_hat); delta = (1-rho_n)/(1+rho_n/2) weights the dipole part."""
    st, stp = np.sqrt(1 - ct * ct), np.sqrt(1 - ctp * ctp)
    th = np.array([ct * np.cos(phi), ct * np.sin(phi), -st])
    ph = np.array([-np.sin(phi), np.cos(phi), 0.0])
    thp = np.array([ctp, 0.0, -stp])
    php = np.array([0.0, 1.0, 0.0])
    a, b, c, d = th @ thp, th @ php, ph @ thp, ph @ php
    m = np.array([[(a * a + b * b + c * c + d * d) / 2, (a * a - b * b + c * c - d * d) / 2, a * b + c * d],
                  [(a * a + b * b - c * c - d * d) / 2, (a * a - b * b - c * c + d * d) / 2, a * b - c * d],
                  [a * c + b * d, a * c - b * d, a * d + b * c]])
    z = 1.5 * delta * m
    z[0, 0] += 1 - delta
    return z


def mode_kernel(m, mu, sgn_out, sgn_in, delta, nphi=64):
    """Kernel of Fourier mode m acting on (I_cos, Q_cos, U_sin), ordered [stokes][node]."""
    n = len(mu)
    k = np.zeros((3, n, 3, n))
    phis = 2 * np.pi * np.arange(nphi) / nphi
    for i in range(n):
        for j in range(n):
            zc, zs = np.zeros((3, 3)), np.zeros((3, 3))
            for p in phis:
                z = mueller_rayleigh(sgn_out * mu[i], p, sgn_in * mu[j], delta)
                zc += z * np.cos(m * p)
                zs += z * np.sin(m * p)
            zc /= nphi
            zs /= nphi
            k[:, i, :, j] = [[zc[0, 0], zc[0, 1], -zs[0, 2]], [zc[1, 0], zc[1, 1], -zs[1, 2]], [zs[2, 0], zs[2, 1], zc[2, 2]]]
    return k.reshape(3 * n, 3 * n)


def vector_doubling(m, mu, w, tau, delta, nd=20):
    n = len(mu)
    d = tau / 2 ** nd
    mu3, w3 = np.tile(mu, 3), np.tile(w, 3)
    c = np.diag(2 * mu3 * w3)
    scale = lambda k: (d / 4 * k / np.outer(mu3, mu3)) @ c
    direct = np.diag(1 - d / mu3)
    # polar axis pointing down: downward travel cos(theta) = +mu, upward = -mu
    r, t = scale(mode_kernel(m, mu, -1, +1, delta)), scale(mode_kernel(m, mu, +1, +1, delta)) + direct
    rs, ts = scale(mode_kernel(m, mu, +1, -1, delta)), scale(mode_kernel(m, mu, -1, -1, delta)) + direct
    eye = np.eye(3 * n)
    for _ in range(nd):
        g1, g2 = np.linalg.inv(eye - r @ rs), np.linalg.inv(eye - rs @ r)
        r, t, rs, ts = r + ts @ g1 @ r @ t, t @ g2 @ t, rs + t @ g2 @ rs @ ts, ts @ g1 @ ts
    return r, t


@pytest.mark.parametrize("ron", [0.0, 0.0279])
def test_polarized_rayleigh_against_vector_doubling(pkg, orc, ron):
    syn = pkg.synth
    ng = 8
    xg, wg = np.polynomial.legendre.leggauss(2 * ng)
    mu, w = xg[ng:][::-1].copy(), wg[ng:][::-1].copy()
    N = ng
    rmu = np.concatenate([-mu[::-1], [0.0], mu])
    ga = np.concatenate([w[::-1], [0.0], w])
    o = syn.make_optics(nb_gauss=8, tetas=40.0, os_nb=16, a_trunc=0.0, piztr=1.0, ipolar=1)
    NT, tau, j0 = 150, 0.5, 3
    h, z = np.linspace(0, tau, NT + 1), np.linspace(100, 0, NT + 1)
    r = orc.sos_os(N, rmu.copy(), ga, o.os_nb, NT, j0, 0.0, 0.0, 0, 0, 1.34, h, np.zeros(NT + 1), np.ones(NT + 1), z, ron,
                   o.alpha.copy(), o.beta, o.gamma.copy(), o.zeta.copy(), -1.0, 100, 2, 1)
    assert r.ier == 0 and r.n_fourier == 3 and r.n_scatter[0] >= 10
    delta = (1 - ron) / (1 + ron / 2)
    scale = np.abs(r.rec[0][2]).max()
    e = np.zeros(3 * N)
    e[j0 - 1] = 1 / (2 * w[j0 - 1])                        # unpolarized unit solar beam at node j0
    for m in (0, 1, 2):
        rr, tt = vector_doubling(m, mu, w, tau, delta)
        up = (rr @ e).reshape(3, N)
        dn = ((tt - np.diag(np.exp(-tau / np.tile(mu, 3)))) @ e).reshape(3, N)
        rec_q, rec_u, rec_i = r.rec[m][0], r.rec[m][1], r.rec[m][2]
        for got, ref in ((rec_i[N + 1:], up[0]), (rec_q[N + 1:], up[1]), (rec_u[N + 1:], -up[2]),
                         (rec_i[:N][::-1], dn[0]), (rec_q[:N][::-1], dn[1]), (rec_u[:N][::-1], -dn[2])):
            assert np.abs(got - ref).max() <= 3e-5 * scale, (m, np.abs(got - ref).max() / scale)
        assert np.abs(up[1]).max() > 1e-3 * scale                  # the polarized components are not trivially small


# ---------------------------------------------------------------------------------------------------------------
# General scattering matrices (aerosols): F(Theta) from the expansion coefficients in the reference's convention,
#   F11 = sum beta_l P_l,  F12 = sum gamma_l sqrt((l-2)!/(l+2)!) P_l^2,
#   F22 +- F33 = sum (alpha_l +- zeta_l) d^l_{2,+-2}   (Wigner d through Jacobi polynomials),
# rotated into the meridian frames with explicit basis vectors.  The construction is calibrated on Rayleigh, where it
# must (and does, to 1e-15) reproduce the dipole phase matrix above.
def _mueller_of_jones(a, b, c, d):
    return np.array([[(a * a + b * b + c * c + d * d) / 2, (a * a - b * b + c * c - d * d) / 2, a * b + c * d],
                     [(a * a + b * b - c * c - d * d) / 2, (a * a - b * b - c * c + d * d) / 2, a * b - c * d],
                     [a * c + b * d, a * c - b * d, a * d + b * c]])


def scattering_matrix(x, al, be, ga, ze):
    from scipy.special import lpmv, gammaln, eval_jacobi, eval_legendre
    f11 = f12 = fp = fm = 0.0
    for l in range(len(be)):
        f11 += be[l] * eval_legendre(l, x)
        if l >= 2:
            f12 += ga[l] * np.exp(0.5 * (gammaln(l - 1) - gammaln(l + 3))) * lpmv(2, l, x)
            fp += (al[l] + ze[l]) / 2 * ((1 + x) / 2) ** 2 * eval_jacobi(l - 2, 0, 4, x)
            fm += (al[l] - ze[l]) / 2 * ((1 - x) / 2) ** 2 * eval_jacobi(l - 2, 4, 0, x)
    return np.array([[f11, f12, 0.0], [f12, fp + fm, 0.0], [0.0, 0.0, fp - fm]])


def phase_matrix(ct, phi, ctp, fmat):
    st, stp = np.sqrt(1 - ct * ct), np.sqrt(1 - ctp * ctp)
    n = np.array([st * np.cos(phi), st * np.sin(phi), ct])
    npr = np.array([stp, 0.0, ctp])
    th = np.array([ct * np.cos(phi), ct * np.sin(phi), -st])
    ph = np.array([-np.sin(phi), np.cos(phi), 0.0])
    thp, php = np.array([ctp, 0.0, -stp]), np.array([0.0, 1.0, 0.0])
    perp = np.cross(npr, n)
    nrm = np.linalg.norm(perp)
    perp = php.copy() if nrm < 1e-12 else perp / nrm          # forward / backward: the meridian plane of n' serves
    parp, par = np.cross(perp, npr), np.cross(perp, n)
    j1 = _mueller_of_jones(parp @ thp, parp @ php, perp @ thp, perp @ php)     # (theta', phi') -> scattering plane
    j2 = _mueller_of_jones(th @ par, th @ perp, ph @ par, ph @ perp)           # scattering plane -> (theta, phi)
    return j2 @ fmat(float(np.clip(n @ npr, -1.0, 1.0))) @ j1


def test_general_phase_matrix_calibrated_on_rayleigh():
    d = 0.9
    al, be, ga, ze = np.array([0, 0, 3 * d]), np.array([1, 0, d / 2]), np.array([0, 0, -np.sqrt(1.5) * d]), np.zeros(3)
    rng = np.random.default_rng(0)
    for _ in range(100):
        ct, ctp = rng.uniform(-1, 1, 2)
        phi = rng.uniform(0, 2 * np.pi)
        z = phase_matrix(ct, phi, ctp, lambda x: scattering_matrix(x, al, be, ga, ze))
        assert np.abs(z - mueller_rayleigh(ct, phi, ctp, d)).max() < 1e-13


def test_polarized_aerosol_layer_against_vector_doubling(pkg, orc):
    """All six kernels of SOS_NOYAUX for l up to OS_NB, the alpha/zeta convention and the sign table of the source
    function, in multiple scattering: aerosol layer (tau 0.4, w0 0.9, Q up to 11 % of I) against the vector doubling."""
    from scipy.interpolate import CubicSpline
    syn = pkg.synth
    ng = 8
    xg, wg = np.polynomial.legendre.leggauss(2 * ng)
    mu, w = xg[ng:][::-1].copy(), wg[ng:][::-1].copy()
    N = ng
    rmu = np.concatenate([-mu[::-1], [0.0], mu])
    gaw = np.concatenate([w[::-1], [0.0], w])
    o = syn.make_optics(nb_gauss=8, tetas=40.0, os_nb=12, a_trunc=0.0, piztr=1.0, ipolar=1, g_modes=((0.6, 1.0),))
    NT, tau, j0, w0 = 150, 0.4, 3, 0.9
    h, z = np.linspace(0, tau, NT + 1), np.linspace(100, 0, NT + 1)
    r = orc.sos_os(N, rmu.copy(), gaw, o.os_nb, NT, j0, 0.0, 0.0, 0, 0, 1.34, h, np.full(NT + 1, w0), np.zeros(NT + 1), z,
                   0.0, o.alpha.copy(), o.beta, o.gamma.copy(), o.zeta.copy(), -1.0, 100, o.os_nb, 1)
    assert r.ier == 0 and r.n_fourier > 4 and r.n_scatter[0] >= 8
    xs = np.linspace(-1, 1, 2001)
    spl = CubicSpline(xs, np.array([scattering_matrix(x, o.alpha, o.beta, o.gamma, o.zeta) for x in xs]).reshape(len(xs), 9))
    fmat = lambda x: spl(x).reshape(3, 3)
    modes, nphi = (0, 1, 2), 32
    phis = 2 * np.pi * np.arange(nphi) / nphi
    kern = {}
    for so, si in ((-1, 1), (1, 1), (1, -1), (-1, -1)):
        ks = {m: np.zeros((3, N, 3, N)) for m in modes}
        for i in range(N):
            for j in range(N):
                zz = np.array([phase_matrix(so * mu[i], p, si * mu[j], fmat) for p in phis])
                for m in modes:
                    zc = (zz * np.cos(m * phis)[:, None, None]).mean(0)
                    zs = (zz * np.sin(m * phis)[:, None, None]).mean(0)
                    ks[m][:, i, :, j] = [[zc[0, 0], zc[0, 1], -zs[0, 2]], [zc[1, 0], zc[1, 1], -zs[1, 2]],
                                         [zs[2, 0], zs[2, 1], zc[2, 2]]]
        kern[(so, si)] = {m: ks[m].reshape(3 * N, 3 * N) for m in modes}
    mu3, w3 = np.tile(mu, 3), np.tile(w, 3)
    c = np.diag(2 * mu3 * w3)
    nd = 20
    d = tau / 2 ** nd
    sc = lambda k: (w0 * d / 4 * k / np.outer(mu3, mu3)) @ c
    scale = np.abs(r.rec[0][2]).max()
    e = np.zeros(3 * N)
    e[j0 - 1] = 1 / (2 * w[j0 - 1])
    for m in modes:
        rr, tt = sc(kern[(-1, 1)][m]), sc(kern[(1, 1)][m]) + np.diag(1 - d / mu3)
        rs, ts = sc(kern[(1, -1)][m]), sc(kern[(-1, -1)][m]) + np.diag(1 - d / mu3)
        eye = np.eye(3 * N)
        for _ in range(nd):
            g1, g2 = np.linalg.inv(eye - rr @ rs), np.linalg.inv(eye - rs @ rr)
            rr, tt, rs, ts = rr + ts @ g1 @ rr @ tt, tt @ g2 @ tt, rs + tt @ g2 @ rs @ ts, ts @ g1 @ ts
        up = (rr @ e).reshape(3, N)
        dn = ((tt - np.diag(np.exp(-tau / mu3))) @ e).reshape(3, N)
        rec_q, rec_u, rec_i = r.rec[m][0], r.rec[m][1], r.rec[m][2]
        for got, ref in ((rec_i[N + 1:], up[0]), (rec_q[N + 1:], up[1]), (rec_u[N + 1:], -up[2]),
                         (rec_i[:N][::-1], dn[0]), (rec_q[:N][::-1], dn[1]), (rec_u[:N][::-1], -dn[2])):
            assert np.abs(got - ref).max() <= 3e-5 * scale, (m, np.abs(got - ref).max() / scale)
    assert np.abs(r.rec[0][0]).max() > 0.05 * scale               # strongly polarized case


def test_polarized_rayleigh_over_lambert_ground(pkg, orc):
    """Vector doubling with a depolarizing Lambertian ground (albedo 0.3, I -> I only, isotropic) coupled in closed form:
    pins the Lambert boundary values (SOS_OS.F:978-980, 1177-1190) in the polarized scattering loop."""
    syn = pkg.synth
    ng = 8
    xg, wg = np.polynomial.legendre.leggauss(2 * ng)
    mu, w = xg[ng:][::-1].copy(), wg[ng:][::-1].copy()
    N = ng
    rmu = np.concatenate([-mu[::-1], [0.0], mu])
    ga = np.concatenate([w[::-1], [0.0], w])
    o = syn.make_optics(nb_gauss=8, tetas=40.0, os_nb=16, a_trunc=0.0, piztr=1.0, ipolar=1)
    NT, tau, j0, rho = 150, 0.5, 3, 0.3
    h, z = np.linspace(0, tau, NT + 1), np.linspace(100, 0, NT + 1)
    r = orc.sos_os(N, rmu.copy(), ga, o.os_nb, NT, j0, 0.0, rho, 0, 0, 1.34, h, np.zeros(NT + 1), np.ones(NT + 1), z, 0.0,
                   o.alpha.copy(), o.beta, o.gamma.copy(), o.zeta.copy(), -1.0, 100, 2, 1)
    mu3 = np.tile(mu, 3)
    n3 = 3 * N
    # all four operators of the layer (illumination from above and from below)
    d = tau / 2 ** 20
    c = np.diag(2 * mu3 * np.tile(w, 3))
    sc = lambda k: (d / 4 * k / np.outer(mu3, mu3)) @ c
    direct = np.diag(1 - d / mu3)
    rr, tt = sc(mode_kernel(0, mu, -1, +1, 1.0)), sc(mode_kernel(0, mu, +1, +1, 1.0)) + direct
    rs, ts = sc(mode_kernel(0, mu, +1, -1, 1.0)), sc(mode_kernel(0, mu, -1, -1, 1.0)) + direct
    eye = np.eye(n3)
    for _ in range(20):
        g1, g2 = np.linalg.inv(eye - rr @ rs), np.linalg.inv(eye - rs @ rr)
        rr, tt, rs, ts = rr + ts @ g1 @ rr @ tt, tt @ g2 @ tt, rs + tt @ g2 @ rs @ ts, ts @ g1 @ ts
    lop = np.zeros((n3, n3))
    lop[:N, :N] = np.tile(2 * rho * mu * w, (N, 1))          # upward I (all nodes) from the downward flux; Q, U -> 0
    e = np.zeros(n3)
    e[j0 - 1] = 1 / (2 * w[j0 - 1])
    dn = np.linalg.solve(eye - rs @ lop, tt @ e)
    up = (rr @ e + ts @ (lop @ dn)).reshape(3, N)
    dnd = (dn - np.exp(-tau / mu3) * e).reshape(3, N)
    scale = np.abs(r.rec[0][2]).max()
    for got, ref in ((r.rec[0][2][N + 1:], up[0]), (r.rec[0][0][N + 1:], up[1]),
                     (r.rec[0][2][:N][::-1], dnd[0]), (r.rec[0][0][:N][::-1], dnd[1])):
        assert np.abs(got - ref).max() <= 3e-5 * scale, np.abs(got - ref).max() / scale


def test_polarized_rayleigh_over_flat_fresnel_sea(pkg, orc):
    """Vector doubling with a flat dielectric interface (n = 1.34): specular reflection with the Fresnel amplitude
    coefficients r_p, r_s in the meridian bases (Jones diag(r_p, r_s) -> Mueller), the specularly reflected direct beam
    removed from the diffuse field as the reference does.  Pins SOS_MAT_FRESNEL_PLAN_REFL, the Fresnel boundary values
    (SOS_OS.F:1225-1239) and SOS_FSOURCE_DIFF_FRESNEL1 for Fourier orders 0-2, I, Q and U."""
    syn = pkg.synth
    ng = 8
    xg, wg = np.polynomial.legendre.leggauss(2 * ng)
    mu, w = xg[ng:][::-1].copy(), wg[ng:][::-1].copy()
    N = ng
    rmu = np.concatenate([-mu[::-1], [0.0], mu])
    ga = np.concatenate([w[::-1], [0.0], w])
    o = syn.make_optics(nb_gauss=8, tetas=40.0, os_nb=16, a_trunc=0.0, piztr=1.0, ipolar=1)
    NT, tau, j0, nref = 150, 0.5, 3, 1.34
    h, z = np.linspace(0, tau, NT + 1), np.linspace(100, 0, NT + 1)
    r = orc.sos_os(N, rmu.copy(), ga, o.os_nb, NT, j0, 0.0, 0.0, 0, 1, nref, h, np.zeros(NT + 1), np.ones(NT + 1), z, 0.0,
                   o.alpha.copy(), o.beta, o.gamma.copy(), o.zeta.copy(), -1.0, 100, 2, 1)
    assert r.ier == 0 and r.n_fourier == 3
    mu3 = np.tile(mu, 3)
    n3 = 3 * N
    lop = np.zeros((n3, n3))
    for i in range(N):
        ci = mu[i]
        ctr = np.sqrt(1 - (1 - ci * ci) / nref ** 2)
        m3 = _mueller_of_jones((nref * ci - ctr) / (nref * ci + ctr), 0.0, 0.0, (ci - nref * ctr) / (ci + nref * ctr))
        for a in range(3):
            for b in range(3):
                lop[a * N + i, b * N + i] = m3[a, b]
    scale = np.abs(r.rec[0][2]).max()
    e = np.zeros(n3)
    e[j0 - 1] = 1 / (2 * w[j0 - 1])
    att = np.diag(np.exp(-tau / mu3))
    for m in (0, 1, 2):
        d = tau / 2 ** 20
        c = np.diag(2 * mu3 * np.tile(w, 3))
        sc = lambda k: (d / 4 * k / np.outer(mu3, mu3)) @ c
        direct = np.diag(1 - d / mu3)
        rr, tt = sc(mode_kernel(m, mu, -1, +1, 1.0)), sc(mode_kernel(m, mu, +1, +1, 1.0)) + direct
        rs, ts = sc(mode_kernel(m, mu, +1, -1, 1.0)), sc(mode_kernel(m, mu, -1, -1, 1.0)) + direct
        eye = np.eye(n3)
        for _ in range(20):
            g1, g2 = np.linalg.inv(eye - rr @ rs), np.linalg.inv(eye - rs @ rr)
            rr, tt, rs, ts = rr + ts @ g1 @ rr @ tt, tt @ g2 @ tt, rs + tt @ g2 @ rs @ ts, ts @ g1 @ ts
        dn = np.linalg.solve(eye - rs @ lop, tt @ e)
        up = ((rr @ e + ts @ (lop @ dn)) - att @ lop @ att @ e).reshape(3, N)   # minus the specular image of the sun
        dnd = (dn - att @ e).reshape(3, N)
        rec_q, rec_u, rec_i = r.rec[m][0], r.rec[m][1], r.rec[m][2]
        for got, ref in ((rec_i[N + 1:], up[0]), (rec_q[N + 1:], up[1]), (rec_u[N + 1:], -up[2]),
                         (rec_i[:N][::-1], dnd[0]), (rec_q[:N][::-1], dnd[1]), (rec_u[:N][::-1], -dnd[2])):
            assert np.abs(got - ref).max() <= 8e-5 * scale, (m, np.abs(got - ref).max() / scale)
