/*
 * sos_oracle.h -- TEST INFRASTRUCTURE ONLY.
 *
 * CPU restatement (plain C, scalar, reference loop order) of the SOS-ABS V5.1
 * successive-orders hot path.  It is the checker for the CUDA product path:
 * only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / reference
 * legs may load it.  The product library (libsosgpu.so) never links or calls it.
 *
 * PARITY PIN: the reference is Fortran 77 and no Fortran compiler exists in this image (nor on the GPU box), and the
 * reference ships no golden vectors for this path.  The pin is therefore the reference's own
 * SOURCE run through a mechanical translation: oracle/build_ref.py + oracle/f77_to_c.py turn SOS.F, SOS_OS.F,
 * SOS_AGGREGATE.F, SOS_TRPHI.F, SOS_GLITTER.F and SOS_SURFACE.F (31 of their 32 subroutines) into C with Fortran's typing
 * rules and gfortran's record I/O, compiled into oracle/_ref/libsosref.so.  tests/test_oracle_vs_reference.py requires
 * this restatement to equal that library BIT FOR BIT (records, loop counts, fluxes, optical depths, surface files, view
 * tables, side effects).  What the pin does not cover is a gfortran BUILD of the same statements (libm is the same glibc;
 * expression evaluation follows the same rules; tests/test_oracle.py bounds the effect of last-bit differences).
 * Independent of any reading of the Fortran, the restatement is also checked against
 * (a) a second, independently written numpy formulation of the order-n source (tests/test_oracle.py),
 * (b) physics: flux conservation, reflection reciprocity per Fourier order, closed-form first-order scattering
 *     (scalar and polarized Rayleigh), scalar and vector adding-doubling solvers for the full multiple-scattering field.
 *
 * Array conventions mirror the reference's Fortran arrays with the *useful*
 * extents instead of the compile-time caps of inc/SOS.h:
 *   angle vectors  V(-N:N)          -> double v[2N+1],      element j at v[j+N]
 *   fields         X(0:NT,-N:N)     -> double x[(2N+1)*(NT+1)], (i,k) at x[(k+N)*(NT+1)+i]
 *   kernels        P(-N:N,-N:N)     -> double p[(2N+1)^2],  (j,k) at p[(k+N)*(2N+1)+(j+N)]
 *   surface record R(1:N,1:N) x 9   -> float  r[9*N*N],     m-th matrix (I,J) at r[m*N*N+(J-1)*N+(I-1)]
 */
#ifndef SOS_ORACLE_H
#define SOS_ORACLE_H

#ifdef __cplusplus
extern "C" {
#endif

/* stop reasons of the scattering loop (SOS_OS.F:1146-1417) */
enum {
  SOS_STOP_IGMAX_PRE = 0, /* IG > IGMAX at label 503 (:1152)            */
  SOS_STOP_GEOM      = 1, /* geometric-series convergence (:1293-1315)  */
  SOS_STOP_LOWVAL    = 2, /* max|X_n| <= 1e-50 (:1368-1382)             */
  SOS_STOP_RATIO     = 3, /* max|X_n/sum| <= 1e-5f (:1387-1402)         */
  SOS_STOP_IGMAX     = 4  /* IG reached IGMAX (:1406-1415)              */
};

/* SOS_NOYAUX (SOS_OS.F:1857-2158).  rmu has 2N+1 entries with rmu[N] = mu_s (index 0). */
void orc_noyaux(int is, int nbmu, const double *rmu, int os_nb,
                const double *alpha, const double *beta, const double *gamma, const double *zeta,
                double *xpl, double *xrl, double *xtl,
                double *bp, double *gr, double *gt, double *arr, double *art, double *att);

/* SOS_FSOURCE_ORDRE1 (SOS_OS.F:2431-2565) */
void orc_fsource_ordre1(int is, int nbmu, int nt, int jk, const double *xdel, const double *ydel,
                        double beta0, double beta2, double gamma2,
                        const double *xpl, const double *xrl, const double *xtl,
                        const double *bp, const double *gr, const double *gt, const double *ch,
                        double *i2, double *q2, double *u2);

/* SOS_FSOURCE_ORDREIG (SOS_OS.F:2663-3017) */
void orc_fsource_ordreig(int is, int nbmu, int nt, const double *xdel, const double *ydel,
                         double beta0, double beta2, double gamma2, double alpha2,
                         const double *xpl, const double *xrl, const double *xtl,
                         const double *i1, const double *q1, const double *u1,
                         const double *bp, const double *gr, const double *gt,
                         const double *arr, const double *art, const double *att, const double *ga,
                         double *i2, double *q2, double *u2);

/* SOS_INTEGR_EPOPT (SOS_OS.F:2222-2357).  i1/q1/u1 carry the ground boundary values for k>0 in. */
void orc_integr_epopt(int nbmu, const double *rmu, int nt, const double *h,
                      const double *i2, const double *q2, const double *u2,
                      double *i1, double *q1, double *u1);

/* SOS_MAT_FRESNEL_PLAN_REFL (SOS_OS.F:1719-1782): f11,f12,f33 have N+1 entries (index 0..N) */
void orc_mat_fresnel_plan_refl(int nbmu, const double *rmu, double ind_surf, int ipolar,
                               double *f11, double *f12, double *f33);

/*
 * SOS_OS (SOS_OS.F:303-1674), in memory.
 *  rmu, ga     : [2N+1], rmu[N] (index 0) is overwritten with mu_s (side effect :715)
 *  alpha..zeta : [os_nb+1], zeroed (alpha,gamma,zeta) when ipolar==0 (side effect :693-697)
 *  surf        : iborm+1 records of 9*N*N REAL*4 (only read when imat_surf==1), may be NULL
 *  rec         : out, (iborm+1) records of 3*(2N+1) doubles in file order Q,U,I (:1572-1574)
 *  n_fourier   : out, number of records written
 *  n_scatter   : out [iborm+1], final IG of each Fourier order
 *  stop_reason : out [iborm+1]
 * returns IER (0 or -1).
 */
int orc_sos_os(int nbmu, double *rmu, const double *ga, int os_nb, int nt,
               int n0, double tetas, double ro, int imat_surf, int ifresnel, double ind_surf,
               const double *h, const double *xdel, const double *ydel, const double *zprof,
               double ron, double *alpha, double *beta, double *gamma, double *zeta,
               double zout, int igmax, int iborm, int ipolar, const float *surf,
               double *rec, int *n_fourier, int *n_scatter, int *stop_reason,
               double *emoins, double *eplus);

/*
 * SOS (SOS.F:340-697), in memory: profile arrays as read from PROFIL_TMP
 * (zprof, h, pcaer, pcmol of length nt+1) are passed in and left untouched; the
 * truncation-adapted copies are returned in h_tr/xdel_tr/ydel_tr (may be NULL).
 * want_trans != 0 reproduces the -SOS.Trans branch (:605-637).
 */
int orc_sos(int nt, double zout, int igmax, int ipolar, double ron, double ind_surf, double rho,
            int imat_surf, int ifresnel, const float *surf, int n0, double piz, double piztr, double a,
            double *rmu, const double *ga, double tetas, int os_nb, int nbmu,
            double *alpha, double *beta, double *gamma, double *zeta,
            const double *zprof, const double *h_in, const double *pcaer, const double *pcmol,
            int want_trans,
            double *rec, int *n_fourier, int *n_scatter, int *stop_reason,
            double *ttot_tronc, double *ttot_vrai, double *tauout, double *tdifmus, double *tdifmug,
            double *emoins, double *eplus, double *h_tr, double *xdel_tr, double *ydel_tr);

/*
 * SOS_AGGREGATE (SOS_AGGREGATE.F:172-543), in memory.
 * res/nres are the running aggregate (the reference's FICOS), tmp/ntmp the term result.
 * have_res==0 reproduces "FICOS does not exist yet".  Returns the new record count, which
 * includes the reference's quirk of appending one all-zero record whenever the existing
 * file is at least as long as the new one.
 */
int orc_aggregate(int nbmu, double aik, const double *tmp, int ntmp,
                  double *res, int nres, int have_res,
                  double ttot_tronc_tmp, double ttot_vrai_tmp, double tauout_tmp,
                  double tdifmus_tmp, const double *tdifmug_tmp, double emoins_tmp, double eplus_tmp,
                  double *ttot_tronc, double *ttot_vrai, double *tauout,
                  double *tdifmus, double *tdifmug, double *emoins, double *eplus);

/* SOS_TRPHI (SOS_TRPHI.F:749-1243); only the glitter / flat-Fresnel direct terms are restated
 * (Roujean / BPDF direct terms are SURVEY 8f "next" rows).  xit/xqt/xut/angdiff: [2N+1]. */
int orc_trphi(const double *rec, int nrec, int nbmu, const double *rmu, double tau, double tauout,
              double phi, int igli, int n0, double wind, double ind_surf, int ifresnel, int ipolar,
              double *xit, double *xqt, double *xut, double *angdiff);

/* SOS_POLAR (SOS_TRPHI.F:1843-1907) */
void orc_polar(double xi, double xq, double xu, double *xan, double *tpol, double *lpol);

/*
 * SOS_TRPHI_OPTION (SOS_TRPHI.F:285-636).  Output tables are [nphi][N] row-major
 * (row = azimuth slot IP, column = J-1) instead of the reference's (0:360,0:80).
 * up/down: 7 tables each in the order SCA, I, Q, U, POL_ANG, POL_RATE, L_POL -> out[7][nphi][N].
 * returns number of azimuth slots filled (2 for itrphi==1) or -1.
 */
int orc_trphi_option(const double *rec, int nrec, int nbmu, const double *rmu, double tau, double tauout,
                     int igli, int n0, double wind, double ind_surf, int ifresnel,
                     int itrphi, double phios, int pas_phi, int ipolar,
                     double *phi_fin, double *theta_fin, double *up, double *down, int nphi_cap);

/* ---- rough-sea reflection matrices (sos_surface_oracle.c) ---- */
/* SOS_GSF for one (theta1, theta2) pair (SOS_GLITTER.F:523-683); e has os_nm+1 entries; returns IL */
int orc_gsf_pair(double c1, double c2, double sig, int os_nm, double *e);
/* SOS_MAT_FRESNEL (SOS_SURFACE.F:1235-1603) incl. the 4(E15.8) round trip of RES_FRESNEL */
void orc_mat_fresnel(int nbmu, const double *rmu, const double *chr, double ind, int os_ns,
                     double *alpha, double *beta, double *gamma, double *zeta);
/* SOS_GLITTER (SOS_GLITTER.F:229): surf [os_nb+1][9][N][N] REAL*4 in the surface-file record layout */
int orc_glitter(int nbmu, const double *rmu, const double *chr, double wind, double ind,
                int os_nb, int os_ns, int os_nm, float *surf, int *il_out);

#ifdef __cplusplus
}
#endif
#endif
