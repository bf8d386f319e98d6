/*
 * sosgpu.h -- C ABI of libsosgpu.so: the B200-native (sm_100a) replacement for the inner loops of
 * SOS-ABS V5.1's successive-orders hot path.
 *
 * Plain C types only (pointers + sizes); every entry point cites the reference interface it replaces.
 * All array arguments are HOST buffers owned by the caller; the library owns device memory, streams
 * and kernels.  There is NO CPU fallback: every compute entry returns SOSGPU_ERR_NO_DEVICE when no
 * CUDA device is usable.
 *
 * Array conventions (useful extents, not the compile-time caps of inc/SOS.h):
 *   angle vectors  V(-N:N)       -> double v[2N+1], element j at v[j+N]
 *   fields         X(0:NT,-N:N)  -> double x[(2N+1)*(NT+1)], (i,k) at x[(k+N)*(NT+1)+i]   (level fastest)
 *   kernels        P(-N:N,-N:N)  -> double p[(2N+1)^2], (j,k) at p[(k+N)*(2N+1)+(j+N)]
 *   surface record               -> float  r[9*N*N], matrix m (R11,R12,R13,R21,..R33), (I,J) at r[m*N*N+(J-1)*N+(I-1)]
 *   Fourier record               -> double rec[3*(2N+1)] in file order Q(-N:N), U(-N:N), I(-N:N)  (SOS_OS.F:1572-1574)
 * The gfortran-ABI shims at the end use the reference's fixed strides instead.
 */
#ifndef SOSGPU_H
#define SOSGPU_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SOSGPU_OK              0
#define SOSGPU_ERR_NO_DEVICE  -2   /* no usable CUDA device: the product path refuses to run */
#define SOSGPU_ERR_CUDA       -3
#define SOSGPU_ERR_ARG        -4
#define SOSGPU_ERR_IER        -1   /* the reference's IER=-1 */

/* compile-time caps mirrored from inc/SOS.h:471,480,202 */
#define SOSGPU_NBMU_MAX 80
#define SOSGPU_NB_MAX   200
#define SOSGPU_NT_MAX   600

/* stop reasons of the scattering loop (SOS_OS.F:1146-1417) */
#define SOSGPU_STOP_IGMAX_PRE 0
#define SOSGPU_STOP_GEOM      1
#define SOSGPU_STOP_LOWVAL    2
#define SOSGPU_STOP_RATIO     3
#define SOSGPU_STOP_IGMAX     4

typedef struct sosgpu_ctx sosgpu_ctx;

/* ---- context ---------------------------------------------------------------------------------- */
int  sosgpu_create(sosgpu_ctx **ctx, int device);
void sosgpu_destroy(sosgpu_ctx *ctx);
const char *sosgpu_last_error(const sosgpu_ctx *ctx);
int  sosgpu_device_count(void);
/* number of kernels this library has launched since create (bench.py's gpu_launches) */
long long sosgpu_launch_count(const sosgpu_ctx *ctx);
/* device time (CUDA events on the library's stream) of the kernel(s) launched by the last sosgpu_glitter,
 * sosgpu_batch_trphi or sosgpu_absprofile / sosgpu_profile / sosgpu_profile_chain call, ms */
double sosgpu_last_kernel_ms(const sosgpu_ctx *ctx);
/* wave sizing knobs: device memory budget for field buffers (bytes, 0 = default) and the
 * maximum number of Fourier orders solved concurrently per term (0 = automatic) */
int  sosgpu_set_options(sosgpu_ctx *ctx, size_t field_budget_bytes, int max_wave_orders);

/* ---- what SOS_PREPA_OS hands to SOS for one wavelength (SOS_PREPA_OS.F:328-335, SOS.F:340-345) */
typedef struct {
  int nbmu;                 /* LUM_NBMU                                                      */
  const double *rmu;        /* [2N+1] cosines, rmu[-j] = -rmu[j]; index 0 is ignored on input */
  const double *ga;         /* [2N+1] weights                                                */
  int n0;                   /* index of the solar angle (>0) or <=0 to use tetas             */
  double tetas;             /* solar zenith angle, degrees                                   */
  int os_nb;                /* OS_NB                                                         */
  const double *alpha, *beta, *gamma, *zeta;   /* [os_nb+1]                                  */
  double a_trunc, piz, piztr;                  /* truncation coefficient, albedos (SOS.F:523-543) */
  double ron;               /* molecular depolarisation factor                               */
  double rho;               /* Lambertian albedo RO                                          */
  int imat_surf, ifresnel;  /* SOS_OS.F:599-612                                              */
  double ind_surf;
  const float *surf;        /* [n_surf_rec][9][N][N] REAL*4 records of the surface file, or NULL */
  int n_surf_rec;
  int igmax, ipolar;
  double zout;              /* -1 = TOA up / BOA down, else altitude (km)                    */
} sosgpu_optics;

/* ---- one (wavelength, CKD term): the PROFIL_TMP content (SOS.F:511-516) + CKD weight ---------- */
typedef struct {
  int optics;               /* index into the optics array                                   */
  int group;                /* aggregation group (wavelength); terms of a group are CKD-summed */
  double aik;               /* CKD weight (SOS_PROC.F:3481-3487)                             */
  int nt;
  const double *zprof, *h, *pcaer, *pcmol;     /* [nt+1] as read from the profile file       */
} sosgpu_term;

/* per-term outputs; arrays sized by the caller with rec_stride = max(os_nb)+1 records */
typedef struct {
  double *rec;              /* [nterm][rec_stride][3][Wmax] zero padded, Wmax = 2*max(nbmu)+1 (may be NULL) */
  int *n_fourier;           /* [nterm]                                                       */
  int *n_scatter;           /* [nterm][rec_stride]                                           */
  int *stop_reason;         /* [nterm][rec_stride]                                           */
  double *emoins, *eplus;   /* [nterm]                                                       */
  double *ttot_tronc, *ttot_vrai, *tauout;     /* [nterm] (SOS.F:518,567-586)                */
  int *ier;                 /* [nterm]                                                       */
} sosgpu_term_out;

/* per-group outputs = what the SOS_AGGREGATE chain leaves behind (SOS_AGGREGATE.F:172-543) */
typedef struct {
  double *rec;              /* [ngroup][rec_stride][3][Wmax] CKD-weighted Fourier coefficients */
  int *n_rec;               /* [ngroup] longest series of the group                          */
  double *emoins, *eplus, *ttot_tronc, *ttot_vrai, *tauout;  /* [ngroup]                     */
} sosgpu_group_out;

/*
 * The batched hot path.  Replaces, for every term, SOS (SOS.F:340) -> SOS_OS (SOS_OS.F:303) and the
 * SOS_AGGREGATE accumulation of SOS_PROC.F:3459-3594, without the file hops.
 * term_out / group_out members may individually be NULL.  rec_stride/wmax describe the caller's layout.
 * When part_only != 0 the group sums are left as partial sums for a multi-GPU reduce
 * (group scalars then hold sum(a*exp(-tau)) instead of -log(...); see sosgpu_group_finalize).
 */
int sosgpu_solve_batch(sosgpu_ctx *ctx, const sosgpu_optics *optics, int noptics,
                       const sosgpu_term *terms, int nterm, int ngroup,
                       int rec_stride, int wmax, int part_only,
                       sosgpu_term_out *term_out, sosgpu_group_out *group_out);

/* turns reduced partial sums (sum a*exp(-tau)) into the reference's -log form (SOS_AGGREGATE.F:467-488) */
int sosgpu_group_finalize(double *ttot_tronc, double *ttot_vrai, double *tauout, int ngroup);

/*
 * Same as sosgpu_solve_batch but keeps inputs resident on the device between calls:
 * upload once, run many (bench.py's HBM-resident timing).  run returns outputs into host buffers
 * only when the out pointers are non-NULL.
 */
typedef struct sosgpu_batch sosgpu_batch;
int  sosgpu_batch_upload(sosgpu_ctx *ctx, const sosgpu_optics *optics, int noptics,
                         const sosgpu_term *terms, int nterm, int ngroup, sosgpu_batch **batch);
/* SOS_OS-level upload: h/pcaer/pcmol are H, XDEL, YDEL exactly as SOS_OS receives them (no truncation
 * adaptation), iborm[nterm] is SOS_OS's IBORM argument (SOS_OS.F:303-308) */
int  sosgpu_batch_upload_os(sosgpu_ctx *ctx, const sosgpu_optics *optics, int noptics,
                            const sosgpu_term *terms, int nterm, int ngroup, const int *iborm,
                            sosgpu_batch **batch);
int  sosgpu_batch_run(sosgpu_ctx *ctx, sosgpu_batch *batch, int rec_stride, int wmax, int part_only,
                      sosgpu_term_out *term_out, sosgpu_group_out *group_out);
void sosgpu_batch_free(sosgpu_ctx *ctx, sosgpu_batch *batch);
/* device pointer + element count of the resident group sums (for an in-place NCCL reduce) */
int  sosgpu_batch_group_buffer(sosgpu_batch *batch, void **dev_ptr, size_t *n_doubles);
/* statistics of the last run: total (term, s, n>=2) contraction steps, their algorithmic FLOPs
 * (2*(6N)^2*(NT+1) each) and recurrence bytes (96*N*(NT+1) each), device ms of the
 * fused step kernel (CUDA events on the launching stream) and its launch count */
typedef struct {
  long long steps; double flops; double bytes; double step_ms; long long step_launches;
  double total_ms; long long launches;
  double aggregate_ms;      /* device ms of the CKD aggregation kernel (SOS_AGGREGATE), CUDA events */
  double useful_flops;      /* the part of `flops` spent on Fourier orders the term ends up keeping (s < n_fourier): a wave
                               solves several orders at once, those past the Fourier stop are computed and discarded */
} sosgpu_stats;
int  sosgpu_batch_stats(const sosgpu_batch *batch, sosgpu_stats *st);

/* ---- multi-GPU: one process per GPU, the library owns the NCCL communicator --------------------------
 * NCCL is bound at run time (dlopen of libnccl.so.2).  Rank 0 creates the unique id, the host program broadcasts its
 * SOSGPU_UNIQUE_ID_BYTES bytes by whatever it has (MPI_Bcast, a file, ...), every rank calls sosgpu_comm_init. */
#define SOSGPU_UNIQUE_ID_BYTES 128
int sosgpu_comm_unique_id(char *id);
int sosgpu_comm_init(sosgpu_ctx *ctx, int nranks, int rank, const char *id);
int sosgpu_comm_destroy(sosgpu_ctx *ctx);
int sosgpu_comm_barrier(sosgpu_ctx *ctx);
/* Term-sharded layout (the CKD terms of every wavelength are spread over the ranks; every rank keeps the global group
 * numbering): after sosgpu_batch_run, ONE in-place ncclReduce (sum, f64) of the partial CKD-weighted sums
 * [ngroup][S+1][3][W] together with the per-group scalars and series lengths forms the band sums on `root`
 * (SOS_AGGREGATE.F:351-488 across GPUs).  Afterwards, on root, sosgpu_batch_trphi synthesises the band sums and
 * sosgpu_batch_groups returns them.  With one rank (or no communicator) it only finalises the local sums. */
int sosgpu_batch_reduce_groups(sosgpu_ctx *ctx, sosgpu_batch *batch, int root);
int sosgpu_batch_groups(sosgpu_ctx *ctx, sosgpu_batch *batch, int rec_stride, int wmax, sosgpu_group_out *group_out);
/* optics entry (index into the uploaded optics array) of every group: needed on a rank that owns no term of a group */
int sosgpu_batch_set_group_optics(sosgpu_batch *batch, const int *optics_of_group);
/* direct[ngroup]: non-zero for a group that is ONE solve whose result the reference does not pass through SOS_AGGREGATE --
 * a wavelength without gaseous absorption or -SOS.AbsModeCKD 2 (SOS_PROC.F:2366, 3609-3716: SOS writes FICSOS_RES_BIN itself and
 * SOS_TRPHI_OPTION gets SOS's own TTOT_TRONC / TAUOUT).  Such a group's optical thicknesses are then the term's own instead of
 * -log(1 * exp(-tau)) (SOS_AGGREGATE.F:467-488), which differs in the last bit.  Ignored for groups of several terms and after
 * sosgpu_batch_reduce_groups (a reduced band sum is an aggregate). */
int sosgpu_batch_set_group_direct(sosgpu_batch *batch, const int *direct);
/* Wavelength-sharded layout (every rank owns whole wavelengths; no reduce): collects the tables of the last
 * sosgpu_batch_trphi call of every rank on `root`, in rank order, with one grouped ncclSend/ncclRecv.
 * groups_of_rank[nranks]; up/down (root; may be NULL): [sum of groups][7][nphi][nmax]. */
int sosgpu_batch_gather_tables(sosgpu_ctx *ctx, sosgpu_batch *batch, int root, const int *groups_of_rank, int nphi, int nmax,
                               double *up, double *down);

/* ---- single-routine operators (parity tests read like the reference's own subroutines) -------- */
/* SOS_NOYAUX (SOS_OS.F:1857-2158): six kernels [W*W] + l=2 rows [W]; rmu[N] must hold mu_s */
int sosgpu_noyaux(sosgpu_ctx *ctx, int is, int nbmu, const double *rmu, int os_nb,
                  const double *alpha, const double *beta, const double *gamma, const double *zeta,
                  double *xpl, double *xrl, double *xtl,
                  double *bp, double *gr, double *gt, double *arr, double *art, double *att);
/* SOS_FSOURCE_ORDREIG (SOS_OS.F:2663-3017) followed by SOS_INTEGR_EPOPT (SOS_OS.F:2222-2357):
 * one fused order step X_{n-1} -> X_n for a single problem over a black surface (zero ground boundary
 * values; bc_reserved must be NULL).  Also returns the source function J (i2,q2,u2) when non-NULL
 * (computed by the same DMMA tiles, staged out for the test). */
int sosgpu_order_step(sosgpu_ctx *ctx, int is, int nbmu, const double *rmu, const double *ga, int os_nb,
                      const double *alpha, const double *beta, const double *gamma, const double *zeta,
                      double ron, int ipolar, int nt, const double *h, const double *xdel, const double *ydel,
                      const double *i1, const double *q1, const double *u1, const double *bc_reserved,
                      double *i1n, double *q1n, double *u1n, double *i2, double *q2, double *u2);

/* Direct-beam terms of the land-surface models in SOS_TRPHI (SOS_TRPHI.F:1047-1200: Roujean BRDF, Rondeaux / Breon /
 * Maignan / Nadal BPDF; helpers SOS_ROUJEAN.F:891-1022, SOS_SURFACE_BPDF.F:1606-1641).  The selection applies to the
 * following sosgpu_trphi_option / sosgpu_batch_trphi calls of the context; NULL switches all of them off. */
typedef struct {
  int iroujean; double k0, k1, k2;
  int irondeaux, ibreon, inadal; double alpha_nadal, beta_nadal;
  int imaignan; double coef_c_maignan;
} sosgpu_direct_models;
int sosgpu_set_direct_models(sosgpu_ctx *ctx, const sosgpu_direct_models *dm);

/* SOS_TRPHI_OPTION / SOS_TRPHI (SOS_TRPHI.F:285-636, 749-1243): Fourier synthesis on the view
 * azimuths + glitter / flat-sea direct terms.  rec: [nrec][3][2N+1].  Tables [7][nphi_cap][N] in the
 * order SCA, I, Q, U, POL_ANG, POL_RATE, L_POL; returns the number of azimuth slots or <0. */
int sosgpu_trphi_option(sosgpu_ctx *ctx, const double *rec, int nrec, int nbmu, const double *rmu,
                        double tau, double tauout, int igli, int n0, double wind, double ind_surf,
                        int ifresnel, int itrphi, double phios, int pas_phi, int ipolar,
                        double *phi_fin, double *theta_fin, double *up, double *down, int nphi_cap);

/* SOS_TRPHI_OPTION for every wavelength (group) of a resident batch, applied to the CKD-summed Fourier coefficients
 * where the last sosgpu_batch_run left them on the device (after an in-place multi-GPU reduce on rank 0).
 * up/down (may be NULL: device-resident timing): [ngroup][7][nphi_cap][Nmax], Nmax = max nbmu of the batch.
 * Returns the number of azimuth slots. */
int sosgpu_batch_trphi(sosgpu_ctx *ctx, sosgpu_batch *batch, int igli, double wind, double ind_surf, int ifresnel,
                       int itrphi, double phios, int pas_phi, int ipolar, int nphi_cap, double *up, double *down);

/* SOS_GLITTER (SOS_GLITTER.F:229) = SOS_GSF + SOS_MAT_FRESNEL + SOS_MAT_REFLEXION + SOS_NOYAUX_FRESNEL +
 * SOS_MISE_FORMAT (SOS_SURFACE.F:1235,1708,2029,2307): Cox-Munk glitter reflection matrices, Fourier-decomposed,
 * in the surface-file record layout.  rmu/chr: [2N+1] cosines / weights; surf: [os_nb+1][9][N][N] REAL*4;
 * il_out (may be NULL): [N(N+1)/2] lengths IL of the G series per (theta1 >= theta2) pair. */
int sosgpu_glitter(sosgpu_ctx *ctx, int nbmu, const double *rmu, const double *chr, int os_nb, int os_ns,
                   int os_nm, double wind, double ind_surf, float *surf, int *il_out);

/* SOS_SURFACE_BPDF (SOS_SURFACE_BPDF.F:219-392) for the Rondeaux (isurf = 4), Breon (5) and Maignan (7, coefficient coef_c)
 * vegetation / soil BPDF models: SOS_GSF_RONDEAUX_BREON or SOS_GSF_MAIGNAN + SOS_MAT_FRESNEL + SOS_MAT_REFLEXION +
 * SOS_MISE_FORMAT; same record layout as sosgpu_glitter.  (Nadal, isurf 6: sosgpu_surface_nadal; here SOSGPU_ERR_ARG.) */
int sosgpu_surface_bpdf(sosgpu_ctx *ctx, int isurf, int nbmu, const double *rmu, const double *chr, int os_nb, int os_ns,
                        int os_nm, double ind_surf, double coef_c, float *surf);

/* SOS_SURFACE_BPDF with ISURF = 6 (SOS_SURFACE_BPDF.F:219-392): Nadal's BPDF model (alpha, beta; > 0) = SOS_F21SF_NADAL +
 * SOS_CALC_F21_NADAL_SUR_FRESNEL (:686-1223: Fourier series of F21 of Nadal / F21 of Fresnel up to os_nb, cut by its recombination
 * test) + SOS_MAT_FRESNEL + SOS_MAT_REFLEXION + SOS_MISE_FORMAT; same record layout as sosgpu_glitter.
 * pairing 0: the reference's file.  SOS_F21SF_NADAL writes a series for each of the N*N (theta1, theta2) pairs and
 * SOS_MAT_REFLEXION reads the series file sequentially for its N(N+1)/2 pairs (I, J <= I) (SOS_SURFACE.F:1832-1842), so that
 * pair number p gets the series of (p / N + 1, p mod N + 1).  pairing 1: every pair gets its own series (not what the reference
 * computes).  il_out (may be NULL): [N(N+1)/2] series lengths IL (-1: the recombination got worse at order 0). */
int sosgpu_surface_nadal(sosgpu_ctx *ctx, int nbmu, const double *rmu, const double *chr, int os_nb, int os_ns, int os_nm,
                         double ind_surf, double alpha, double beta, int pairing, float *surf, int *il_out);

/* SOS_ROUJEAN (SOS_ROUJEAN.F:212-416: SOS_FSF_ROUJEAN + SOS_MISE_FORMAT_RJ): Fourier series of Roujean's BRDF (k0, k1, k2) for
 * every (incidence, reflection) pair, surf [os_nb+1][9][N][N] REAL*4 (only R11 non-zero).  SOSGPU_ERR_IER when the model
 * gives a negative BRDF (the reference's label 993). */
int sosgpu_roujean(sosgpu_ctx *ctx, int nbmu, const double *rmu, int os_nb, double k0, double k1, double k2, float *surf);
/* SOS_BPDF_AJOUT_BRDF (SOS_SURFACE.F:2503-2669): out = surf1 + surf2, record by record (BPDF + BRDF land surface) */
int sosgpu_bpdf_ajout_brdf(sosgpu_ctx *ctx, const float *surf1, const float *surf2, int nbmu, int os_nb, float *out);

/* SOS_MAT_FRESNEL (SOS_SURFACE.F:1235-1603): Legendre expansion of the Fresnel reflection matrix for the refractive index
 * ind_surf on the quadrature (rmu, chr: [2N+1]); returns what the reference's RES_FRESNEL file holds, i.e. the coefficients
 * [os_ns+1] after the 4(E15.8) decimal round trip. */
int sosgpu_mat_fresnel(sosgpu_ctx *ctx, int nbmu, const double *rmu, const double *chr, double ind_surf, int os_ns,
                       double *alpha, double *beta, double *gamma, double *zeta);

/* SOS_Up.txt / SOS_Down.txt (SOS_ABS_MAIN.F:2250-2519, record formats :3095-3096, headers SOS_TRPHI.F:1570-1796) from the
 * tables of sosgpu_trphi_option ([7][nphi_cap][N]: SCA, I, Q, U, POL_ANG, POL_RATE, L_POL), byte-compatible with the
 * reference.  fix_sca_index = 0 keeps the reference's indexing of the upward scattering angle by degrees in view mode 2
 * (SOS_ABS_MAIN.F:2467); host-only (no device needed). */
int sosgpu_write_updown(const char *fic_up, const char *fic_down, int nbmu, int itrphi, double phios, int pas_phi, double zout,
                        const double *phi_fin, const double *theta_fin, const double *up, const double *down, int nphi_cap,
                        int fix_sca_index);

/* The optional transmission and flux files SOS_PROC writes after the synthesis (SOS_PROC.F:3779-3874, FORMAT :4944-4951); host-only.
 * rmu / tdifmug: RMU(1:N), TDIFMUG(1:N).  zalt: ABS_USERPROFIL(1:50,1); tauabs: TAUABS(1:50) of the last term, as there.
 * sosgpu_write_flux also returns TDIR_VRAI, FLUX_DIFF_DOWN, FLUX_DOWN (SOS_PROC outputs) and writes nothing for 'NO_OUTPUT'.
 * The list-directed records follow gfortran's documented layout (unpinned: no Fortran compiler in the build environment). */
int sosgpu_write_trans(const char *fictrans, double tetas, double ttot_tronc, double ttot_vrai, double tdifmus, int nbmu,
                       const double *rmu, const double *tdifmug);
int sosgpu_write_flux(const char *ficflux, double tetas, double ttot_tronc, double ttot_vrai, double emoins, double eplus,
                      double tr, double hr, double ta, double ha, const double *zalt, const double *tauabs,
                      double *tdir_vrai, double *flux_diff_down, double *flux_down);

/* ---- the per-term profile chain that precedes every term-solve (SOS_PROC.F:3459-3537) ------------
 * For each (wavelength, CKD term): SOS_ABSPROFILE (SOS_ABSPROFILE.F:184-425, with COEFF_ABS_CKD SOS_SUB_TRS.F:171-393)
 * gives the gas optical thickness at the 50 levels of the gas atmosphere, SOS_PROFILE (SOS_PROFIL.F:224-1156, SOS_DISC :1210)
 * cuts the atmosphere into NT layers of about equal optical thickness, and the PROFIL_TMP text file (format 2X,I5,F10.5,
 * 3(E15.8)) carries ZPROF, H, PCAER, PCMOL to SOS (SOS.F:511-516).  Here the chain runs as device kernels over all terms of a
 * band at once and its outputs are the arrays sosgpu_term points to.  Tables in the reference's Fortran storage (column-major,
 * the compile-time extents of inc/SOS.h:246-282), as READ_CKD_COEFF and SOS_PREPA_ABSPROFILE fill them. */
typedef struct {
  int nb_temp, nb_pres, nb_conc_h2o;
  const double *tab_temp;        /* TAB_TEMP(9)                    */
  const double *tab_pres;        /* TAB_PRES(31)                   */
  const double *tab_conc_h2o;    /* TAB_CONC_H2O(12)               */
  const int    *nexp;            /* NEXP(8,50)                     */
  const double *kdis_ki;         /* KDIS_KI(9,31,5,8,50)           */
  const double *kdis_ki_h2o;     /* KDIS_KI_H2O(9,31,12,5,50)      */
} sosgpu_ckd;
typedef struct {
  const double *userprofil;      /* USERPROFIL(50,13): altitude, P (mbar), T (K), H2O ... (SOS_ABSPROFILE.F:91-104) */
  const double *altabs;          /* ALTABS(50), descending         */
  const double *ro;              /* RO(8,50) particles/cm2 per gas and layer */
} sosgpu_gas_profile;
typedef struct {
  int lamb1;                     /* spectral interval of the CKD tables, 1-based (unused when absprofil = 7) */
  int ik[8];                     /* IK1..IK8: exponential index per gas (H2O CO2 O3 N2O CO CH4 O2 NO2) */
  int absprofil;                 /* 7: no gaseous absorption       */
  int iprofil;                   /* 1: exponential profiles, 2: aerosols between zmin and zmax */
  double tr, hr, ta, ha, zmin, zmax;
} sosgpu_profile_term;
/* per-term error codes returned in ier[]: 0, the reference's error label (900..923 COEFF_ABS_CKD, 940 / 1010 / 1020
 * SOS_PROFILE), or 9600 when a profile would need more than SOSGPU_NT_MAX levels (the reference writes out of bounds there) */
/* SOS_ABSPROFILE for nterm terms: tauabs [nterm][50] */
int sosgpu_absprofile(sosgpu_ctx *ctx, const sosgpu_ckd *ckd, const sosgpu_gas_profile *atm, const sosgpu_profile_term *terms,
                      int nterm, double *tauabs, int *ier);
/* SOS_PROFILE for nterm terms from given absorption profiles tauabs [nterm][50]; outputs nt [nterm] and zprof, h, pcaer,
 * pcmol [nterm][SOSGPU_NT_MAX+1].  text_hop = 1 returns the values SOS reads back from PROFIL_TMP (5 decimals / 8 significant
 * digits), 0 the unrounded ones */
int sosgpu_profile(sosgpu_ctx *ctx, const double *altabs, const double *tauabs, const sosgpu_profile_term *terms, int nterm,
                   int text_hop, int *nt, double *zprof, double *h, double *pcaer, double *pcmol, int *ier);
/* both stages without the host hop in between; tauabs may be NULL */
int sosgpu_profile_chain(sosgpu_ctx *ctx, const sosgpu_ckd *ckd, const sosgpu_gas_profile *atm, const sosgpu_profile_term *terms,
                         int nterm, int text_hop, double *tauabs, int *nt, double *zprof, double *h, double *pcaer, double *pcmol,
                         int *ier);
/* READ_CKD_COEFF (SOS_SUB_TRS.F:481-905): host-side reader of $SOS_ABS_ROOT/fic/COEFF_CKD/<step>cmm1/coef_<GAS>_<max>_<min>_
 * <step>cmm1 for gas nabs (1..8) and wavenumber nu; jabs = 0 fills the "gas not selected" defaults.  sos_abs_root NULL reads the
 * environment variable as the reference does.  kdis_ai is KDIS_AI(5,8,50).  Returns 0 or -1 (the reference's IER). */
int sosgpu_read_ckd_coeff(const char *sos_abs_root, int nabs, int jabs, double nu, double nustep, int *nexp, double *kdis_ai,
                          double *kdis_ki, double *kdis_ki_h2o, double *numax, double *numin, double *tab_pres, int *nb_pres,
                          double *tab_temp, int *nb_temp, double *tab_conc_h2o, int *nb_conc_h2o);

/* ---- aerosol optics per wavelength ----------------------------------------------------------------
 * What SOS_AEROSOLS does on the host before a wavelength is solved: Mie theory on a grid of size parameters (SOS_MIE,
 * SOS_FPHASE_MIE, SOS_MIE.F:205-690, 801-944; cached in MIE files keyed by refractive index and size-parameter range), the
 * integral over a size distribution (SOS_GRANU, SOS_AEROSOLS.F:4392-4767), the mixture of modes (SOS_AEROSOLS.F:1455-1490,
 * 2085-2110) and the Legendre expansion with the optional truncation of the forward peak (SOS_DECOMPO_LEGENDRE,
 * SOS_AEROSOLS.F:3924-4260), whose ALPHA, BETA11, GAMMA12, ZETA (0:os_nb) and truncated albedo are the solver's inputs.
 * Angle vectors V(-nbmu:nbmu) as everywhere: double v[2*nbmu+1], element j at v[j+nbmu]; nang = 2*nbmu+1.
 * A Mie table has one record per size parameter: rec[k][3] = ALPHA, QEXT, QSCA (REAL*4), g[k] (REAL*8) and the phase
 * functions imie / qmie / umie [k][nang] (REAL*4) -- the contents of the reference's MIE file. */
#define SOSGPU_MIE_NBMU_MAX 100     /* CTE_MIE_NBMU_MAX (inc/SOS.h:457) */
#define SOSGPU_MIE_DIM      10000   /* CTE_MIE_DIM (inc/SOS.h:96): alphaf + alphaf + 20 must not exceed it */
typedef struct {
  double rn, in;             /* refractive index of the particles; in <= 0 */
  double alpha0, alphaf;     /* size-parameter range of the Mie table (SOS_MIE's ALPHAO, ALPHAF) */
  int igranu;                /* 1: log-normal, v1 = modal radius (microns), v2 = sigma; 2: Junge, v1 = r0, v2 = slope, v3 = rmax */
  double v1, v2, v3;
  double wa;                 /* wavelength (microns) */
} sosgpu_aer_component;
typedef struct {
  int ncomp;                 /* 0: component comp[0] as it is (mono-modal); 1..4: mixture of comp[0..ncomp-1] */
  int comp[4];               /* indices into the component list */
  double weight[4];          /* number fractions: N(I)/NTOT (WMO) or the normalised CVI(I) (bimodal); 0 skips the component */
  int itronc;                /* 1: truncate the forward peak (cancelled when the coefficient stays below 0.1) */
} sosgpu_aer_model;
/* number of records SOS_MIE writes for [alpha0, alphaf] (variable step, SOS_MIE.F:406-411), or -1 (its error 997) */
int sosgpu_mie_count(double alpha0, double alphaf);
/* SOS_MIE for one table; outputs are host arrays of capacity nrec_cap records, *nrec the number written */
int sosgpu_mie(sosgpu_ctx *ctx, int nbmu, const double *rmu, double rn, double in, double alpha0, double alphaf, int nrec_cap,
               float *rec, double *g, float *imie, float *qmie, float *umie, int *nrec);
/* SOS_GRANU on a Mie table in host arrays (alphaf: the table's header value).  kmat[3] = KMAT1, KMAT2 (cross sections,
 * square microns), SOMME_NR; *ier = -1 when the table ends before the size distribution does (the reference's read error) */
int sosgpu_granu(sosgpu_ctx *ctx, int nbmu, int nrec, const float *rec, const float *imie, const float *qmie, const float *umie,
                 double alphaf, int igranu, double v1, double v2, double v3, double wa, double *kmat, double *p11, double *p12,
                 double *p33, int *ier);
/* SOS_DECOMPO_LEGENDRE: p11 in / out (truncated on exit), ttt out (P11 before truncation), *itronc in / out; the coefficient
 * arrays (0:os_nb) are overwritten (the reference accumulates into arrays its caller has zeroed) */
int sosgpu_decompo_legendre(sosgpu_ctx *ctx, int *itronc, int nbmu, const double *xmu, const double *xhr, int os_nb, double *p11,
                            double *ttt, const double *p12, const double *p22, const double *p33, double *coef_tronca, double *z1,
                            double *alp, double *beta11, double *beta22, double *gamma12, double *delta33, double *zeta, int *ier);
/* The whole chain on the device for ncomp components (one per mode and wavelength) and nmodel models (one per wavelength):
 * one Mie table per distinct (rn, in, alpha0, alphaf), never copied to the host.  Host outputs (NULL = not wanted):
 *   comp_k [ncomp][3] KMAT1, KMAT2, SOMME_NR; comp_phase [ncomp][3][nang] P11, P12, P33; comp_ier [ncomp];
 *   scal [nmodel][8] KMAT1, KMAT2, PIZ, PIZTR (albedo after truncation), COEF_TRONCA, asymmetry factor, Z1, ITRONC on exit;
 *   coef [nmodel][6][os_nb+1] ALPHA, BETA11, GAMMA12, ZETA, BETA22, DELTA33; phase [nmodel][4][nang] P11 (truncated), P12,
 *   P33, TTT; model_ier [nmodel] */
int sosgpu_aerosols(sosgpu_ctx *ctx, int nbmu, const double *xmu, const double *xhr, int ncomp, const sosgpu_aer_component *comp,
                    int nmodel, const sosgpu_aer_model *models, int os_nb, double *comp_k, double *comp_phase, int *comp_ier,
                    double *scal, double *coef, double *phase, int *model_ier);
/* the aerosol result file SOS_AEROSOLS writes and SOS_PREPA_OS reads (SOS_AEROSOLS.F:2810-2832, formats 39-50); host only */
int sosgpu_write_aerosols(const char *path, int os_nb, double kmat1, double kmat2, double asym, double coef_tronca, double piztr,
                          const double *alp, const double *beta11, const double *gamma12, const double *zeta);

/* ---- gfortran-ABI drop-in symbols (F77 by-reference, fixed SOS.h strides, hidden string lengths) */
/* SOS_OS.F:303-308 */
void sos_os_(const int *nbmu, double *rmu, const double *ga, const int *os_nb, const int *nt,
             const char *ficsurf, const char *ficos,
             const int *n0, const double *tetas, const double *ro, const int *imat_surf,
             const int *ifresnel, const double *ind_surf,
             const double *h, const double *xdel, const double *ydel, const double *zprof, const double *ron,
             double *alpha, double *beta, double *gamma, double *zeta, const double *zout,
             const int *igmax, const int *iborm, const int *ipolar, const int *trace, const int *idlog,
             double *emoins, double *eplus, int *ier, size_t len_ficsurf, size_t len_ficos);
/* SOS_AGGREGATE.F:172-178 */
void sos_aggregate_(const int *nbmu, const double *aik, const char *ficos_tmp,
                    const double *ttot_tronc_tmp, const double *ttot_vrai_tmp, const double *tauout_tmp,
                    const double *tdifmus_tmp, const double *tdifmug_tmp, const double *emoins_tmp,
                    const double *eplus_tmp, const char *ficos_agg_tmp, const char *ficos,
                    double *ttot_tronc, double *ttot_vrai, double *tauout, double *tdifmus, double *tdifmug,
                    double *emoins, double *eplus, int *ier,
                    size_t len_ficos_tmp, size_t len_ficos_agg_tmp, size_t len_ficos);

/* SOS.F:340-345 (hidden lengths in order of appearance: FICOS, FICTRANS, FICPROFIL, FICSURF).  Reads FICPROFIL,
 * writes FICOS; FICTRANS is only compared with 'NO_OUTPUT' (the transmissions come back in TDIFMUS/TDIFMUG) */
void sos_(const char *ficos, const char *fictrans, const char *ficprofil,
          const int *nt, const double *zout, const int *igmax, const int *ipolar, const double *ron,
          const double *ind_surf, const double *rho, const int *imat_surf, const int *ifresnel,
          const char *ficsurf, const int *n0, const double *piz, const double *piztr, const double *a,
          double *rmu, const double *ga, const double *tetas, const int *os_nb, const int *lum_nbmu,
          double *alpha, double *beta, double *gamma, double *zeta,
          double *ttot_tronc, double *ttot_vrai, double *tauout, double *tdifmus, double *tdifmug,
          double *emoins, double *eplus, const int *trace, const int *idlog, int *ier,
          size_t len_ficos, size_t len_fictrans, size_t len_ficprofil, size_t len_ficsurf);
/* SOS_GLITTER.F:229-233: writes FICGLITTER (must not pre-exist); the three intermediate files are never created */
void sos_glitter_(const int *lum_nbmu, const double *rmu, const double *chr, const double *wind,
                  const double *ind, const int *os_nb, const int *os_ns, const int *os_nm,
                  const char *fic_res_gsf, const char *fic_res_fresnel, const char *fic_res_mat_reflex,
                  const char *ficglitter, const int *trace, int *ier,
                  size_t len_gsf, size_t len_fresnel, size_t len_mat, size_t len_ficglitter);
/* SOS_TRPHI.F:749-755: one azimuth (radians), incl. the Roujean / Rondeaux / Breon / Nadal / Maignan direct terms */
void sos_trphi_(const char *fichos, const int *nbmu, const double *rmu, const double *tau,
                const double *tauout, const double *phi, const int *igli, const int *n0, const double *wind,
                const double *ind_surf, const int *ifresnel, const int *iroujean, const double *k0,
                const double *k1, const double *k2, const int *irondeaux, const int *ibreon,
                const int *inadal, const double *alpha_nadal, const double *beta_nadal, const int *imaignan,
                const double *coef_c_maignan, const int *ipolar, double *xit, double *xqt, double *xut,
                double *angdiff, int *ier, size_t len_fichos);
/* SOS_TRPHI.F:285-300: tables with the reference's extents PHI_FIN(0:360), THETA_FIN(0:80), X_FIN(0:360,0:80) */
void sos_trphi_option_(const int *nbmu, const double *rmu, const double *ga, const char *fichos,
                       const double *tau, const double *tauout, const double *zout, const int *igli,
                       const int *n0, const double *wind, const double *ind_surf, const int *ifresnel,
                       const int *iroujean, const double *k0, const double *k1, const double *k2,
                       const int *irondeaux, const int *ibreon, const int *inadal, const double *alpha_nadal,
                       const double *beta_nadal, const int *imaignan, const double *coef_c_maignan,
                       const int *itrphi, const double *phios, const int *pas_phi, const int *ipolar,
                       double *phi_fin, double *theta_fin,
                       double *sca_up, double *i_up, double *q_up, double *u_up, double *pol_ang_up,
                       double *pol_rate_up, double *l_pol_up,
                       double *sca_down, double *i_down, double *q_down, double *u_down, double *pol_ang_down,
                       double *pol_rate_down, double *l_pol_down, int *ier, size_t len_fichos);

/* SOS_SUB_TRS.F:481-485 (INTEGER*2 arguments are short) */
void read_ckd_coeff_(const short *nabs, const short *jabs, const double *nu, const double *nustep, int *nexp, double *kdis_ai,
                     double *kdis_ki, double *kdis_ki_h2o, double *numax, double *numin, double *tab_pres, int *nb_pres,
                     double *tab_temp, int *nb_temp, double *tab_conc_h2o, int *nb_conc_h2o, int *ier);
/* SOS_ABSPROFILE.F:184-190 */
void sos_absprofile_(const short *absprofil, const double *nu, const int *lamb1, const short *iabs, const double *userprofil,
                     const double *altabs, const double *ro, const int *nexp, const double *kdis_ki, const double *kdis_ki_h2o,
                     const int *ik1, const int *ik2, const int *ik3, const int *ik4, const int *ik5, const int *ik6,
                     const int *ik7, const int *ik8, const double *tab_pres, const int *nb_pres, const double *tab_temp,
                     const int *nb_temp, const double *tab_conc_h2o, const int *nb_conc_h2o, double *tauabstot,
                     const int *trace, const int *idlog, int *ier);
/* SOS_PROFIL.F:224-226: writes FICPROFIL (format 20), returns NT */
void sos_profile_(const short *iprofil, const double *tr, const double *hr, const double *ta, const double *ha,
                  const double *zmin, const double *zmax, const short *absprofil, const double *altabs, const double *tabs,
                  const int *trace, const int *idlog, const char *ficprofil, int *nt, int *ier, size_t len_ficprofil);

/* SOS_MIE.F:205-206: writes the unformatted MIE file (hidden lengths of FICMIE, FICLOG last); angle vectors (-100:100) */
void sos_mie_(const int *mie_nbmu, const double *rmu, const double *chr, const double *rn, const double *in, const double *alphao,
              const double *alphaf, const char *ficmie, const char *ficlog, int *ier, size_t len_ficmie, size_t len_ficlog);
/* SOS_AEROSOLS.F:4392-4394: reads the MIE file */
void sos_granu_(const char *ficmie, const int *igranu, const double *v1, const double *v2, const double *v3, const double *wa,
                const int *mie_nbmu, const double *xmu, const int *trace, double *kmat1, double *kmat2, double *somme_nr,
                double *p11, double *p12, double *p33, int *ier, size_t len_ficmie);
/* SOS_AEROSOLS.F:3924-3928: coefficient arrays (0:200) */
void sos_decompo_legendre_(int *itronc, const int *trace, const int *mie_nbmu, const double *xmu, const double *xhr, const int *os_nb,
                           double *p11, double *ttt, const double *p12, const double *p22, const double *p33, double *coef_tronca,
                           double *z1, double *alp, double *beta11, double *beta22, double *gamma12, double *delta33, double *zeta,
                           int *ier);

#ifdef __cplusplus
}
#endif
#endif
