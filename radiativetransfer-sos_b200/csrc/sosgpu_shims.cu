// sosgpu_shims.cu -- (1) azimuth synthesis entry points (SOS_TRPHI_OPTION), single wavelength and batched over the
// resident group sums; (2) gfortran-ABI drop-in symbols sos_ / sos_os_ / sos_aggregate_ / sos_glitter_ / sos_trphi_ /
// sos_trphi_option_ that keep the reference's argument lists, fixed SOS.h strides and file side effects
// (SOS.F:340-345, SOS_OS.F:303-308, SOS_AGGREGATE.F:172-178, SOS_GLITTER.F:229-233, SOS_TRPHI.F:285-300,749-755).
#include "sosgpu_host.h"
#include "post_kernels.h"

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

// ---------------------------------------------------------------------------------------------
static int phis_of(int itrphi, double phios, int pas_phi, std::vector<double> &phis, std::vector<double> &phi_fin)
{
  const double pi = std::acos(-1.0);
  phis.clear(); phi_fin.clear();
  if (itrphi == 1) {                                           // SOS_TRPHI.F:435,494
    phis.push_back(pi + phios * pi / 180.0);
    phis.push_back(phios * pi / 180.0);
    phi_fin.push_back(phios);                                  // PHI_FIN(0) ends up holding PHIOS (:445,504)
    phi_fin.push_back(0.0);
  } else if (itrphi == 2) {                                    // SOS_TRPHI.F:558-561
    if (pas_phi < 1) return -1;
    for (int iphi = 0; iphi <= 360; iphi += pas_phi) {
      phis.push_back(pi * iphi / 180.0);
      phi_fin.push_back((double)iphi);
    }
  } else return -1;
  return (int)phis.size();
}

// Synthesis of one wavelength on a list of azimuths (radians): out = [2 (up, down)][7][nphi][N]
static int trphi_core(sosgpu_ctx *ctx, const double *rec, int nrec, int N, const double *rmu, double tau, double tauout,
                      int igli, int n0, double wind, double ind_surf, int ifresnel, int ipolar,
                      const std::vector<double> &phis, std::vector<double> &out)
{
  CK(cudaSetDevice(ctx->device));
  const int W = 2 * N + 1, nphi = (int)phis.size();
  double *d_rec = nullptr, *d_rmu = nullptr, *d_phi = nullptr, *d_out = nullptr;
  TrphiGroup *d_g = nullptr;
  const size_t nout = (size_t)2 * 7 * nphi * N;
  SosFreeGuard guard(ctx);                                       // pool allocations, released on every return path
  CK(sos_dmalloc(ctx, &d_rec, (size_t)nrec * 3 * W * 8)); guard.add(d_rec);
  CK(sos_dmalloc(ctx, &d_rmu, W * 8)); guard.add(d_rmu);
  CK(sos_dmalloc(ctx, &d_phi, nphi * 8)); guard.add(d_phi);
  CK(sos_dmalloc(ctx, &d_out, nout * 8)); guard.add(d_out);
  CK(sos_dmalloc(ctx, &d_g, sizeof(TrphiGroup))); guard.add(d_g);
  CK(cudaMemcpyAsync(d_rec, rec, (size_t)nrec * 3 * W * 8, cudaMemcpyHostToDevice, ctx->stream));
  CK(cudaMemcpyAsync(d_rmu, rmu, W * 8, cudaMemcpyHostToDevice, ctx->stream));
  CK(cudaMemcpyAsync(d_phi, phis.data(), nphi * 8, cudaMemcpyHostToDevice, ctx->stream));
  TrphiGroup g{d_rec, d_rmu, nrec, N, n0, W, tau, tauout};
  CK(cudaMemcpyAsync(d_g, &g, sizeof(g), cudaMemcpyHostToDevice, ctx->stream));
  const TrphiParams prm = sos_trphi_params(ctx, igli, ifresnel, ipolar, wind, ind_surf, std::acos(-1.0));
  sos_launch_trphi(d_g, 1, d_phi, nphi, prm, d_out, ctx->stream);
  ctx->launches += 1;
  out.resize(nout);
  CK(cudaMemcpyAsync(out.data(), d_out, nout * 8, cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  CK(cudaGetLastError());
  return SOSGPU_OK;
}

extern "C" int sosgpu_trphi_option(sosgpu_ctx *ctx, const double *rec, int nrec, int nbmu, const double *rmu,
                                   double tau, double tauout, int igli, int n0, double wind, double ind_surf,
                                   int ifresnel, int itrphi, double phios, int pas_phi, int ipolar,
                                   double *phi_fin, double *theta_fin, double *up, double *down, int nphi_cap)
{
  if (!ctx) return SOSGPU_ERR_NO_DEVICE;
  if (!rec || nrec < 1 || nbmu < 1 || nbmu > SOSGPU_NBMU_MAX || n0 < 1 || n0 > nbmu) return SOSGPU_ERR_ARG;
  const int N = nbmu;
  std::vector<double> phis, pf, out;
  const int nphi = phis_of(itrphi, phios, pas_phi, phis, pf);
  if (nphi < 1 || nphi > nphi_cap) return SOSGPU_ERR_ARG;
  const int rc = trphi_core(ctx, rec, nrec, N, rmu, tau, tauout, igli, n0, wind, ind_surf, ifresnel, ipolar, phis, out);
  if (rc != SOSGPU_OK) return rc;
  // [2][7][nphi][N] -> caller tables [7][nphi_cap][N]
  for (int ud = 0; ud < 2; ++ud) {
    double *dst = ud == 0 ? up : down;
    if (!dst) continue;
    for (int t = 0; t < 7; ++t)
      for (int ip = 0; ip < nphi; ++ip)
        memcpy(dst + ((size_t)t * nphi_cap + ip) * N, &out[(((size_t)ud * 7 + t) * nphi + ip) * N], N * 8);
  }
  const double pi = std::acos(-1.0);
  if (phi_fin) for (int ip = 0; ip < nphi; ++ip) phi_fin[ip] = pf[ip];
  if (theta_fin) for (int j = 1; j <= N; ++j) theta_fin[j - 1] = std::acos(rmu[j + N]) * 180.0 / pi;   // :508,573
  return nphi;
}

// ---------------------------------------------------------------------------------------------
// SOS_MAT_FRESNEL (SOS_SURFACE.F:1235-1603) on the host: O(N*NS + NS^2) scalar work whose output goes through the
// reference's 4(E15.8) text file (RES_FRESNEL, :1552 / :1822) -- the decimal round trip is a host operation.
static double round_e15_8(double x)
{
  char buf[64];
  snprintf(buf, sizeof buf, "%.7E", x);
  return strtod(buf, nullptr);
}
// SOS_MAT_FRESNEL (SOS_SURFACE.F:1235-1603): the expansion runs on the device (k_mat_fresnel, glitter_kernel.cu); its result
// reaches SOS_MAT_REFLEXION through the reference's 4(E15.8) text file (RES_FRESNEL, :1552 / :1822), i.e. rounded to 8
// significant decimal digits -- that decimal round trip is the host's part.  d_in: [W rmu | W chr | 4*(ns+1) coefficients].
static int mat_fresnel_device(sosgpu_ctx *ctx, int N, double ind, int ns, double *d_in, std::vector<double> &coef)
{
  const int W = 2 * N + 1;
  double *d_coef = d_in + 2 * W;
  sos_launch_mat_fresnel(N, d_in, d_in + W, ind, ns, d_coef, ctx->stream);
  ctx->launches += 1;
  coef.resize((size_t)4 * (ns + 1));
  CK(cudaMemcpyAsync(coef.data(), d_coef, coef.size() * 8, cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  CK(cudaGetLastError());
  for (double &v : coef) v = round_e15_8(v);
  CK(cudaMemcpyAsync(d_coef, coef.data(), coef.size() * 8, cudaMemcpyHostToDevice, ctx->stream));
  return SOSGPU_OK;
}

extern "C" int sosgpu_mat_fresnel(sosgpu_ctx *ctx, int nbmu, const double *rmu, const double *chr, double ind_surf, int os_ns,
                                  double *alpha, double *beta, double *gamma, double *zeta)
{
  if (!ctx) return SOSGPU_ERR_NO_DEVICE;
  if (!rmu || !chr || nbmu < 1 || nbmu > SOSGPU_NBMU_MAX || os_ns < 2 || os_ns > 136) { ctx->err = "sosgpu_mat_fresnel: bad arguments"; return SOSGPU_ERR_ARG; }
  CK(cudaSetDevice(ctx->device));
  const int W = 2 * nbmu + 1;
  double *d_in = nullptr;
  CK(sos_dmalloc(ctx, &d_in, (size_t)(2 * W + 4 * (os_ns + 1)) * 8));
  struct Guard { sosgpu_ctx *c; void *p; ~Guard() { sos_dfree(c, p); } } guard{ctx, d_in};
  CK(cudaMemcpyAsync(d_in, rmu, W * 8, cudaMemcpyHostToDevice, ctx->stream));
  CK(cudaMemcpyAsync(d_in + W, chr, W * 8, cudaMemcpyHostToDevice, ctx->stream));
  std::vector<double> coef;
  const int rc = mat_fresnel_device(ctx, nbmu, ind_surf, os_ns, d_in, coef);
  if (rc != SOSGPU_OK) return rc;
  CK(cudaStreamSynchronize(ctx->stream));
  const size_t n = os_ns + 1;
  if (alpha) memcpy(alpha, &coef[0], n * 8);
  if (beta) memcpy(beta, &coef[n], n * 8);
  if (gamma) memcpy(gamma, &coef[2 * n], n * 8);
  if (zeta) memcpy(zeta, &coef[3 * n], n * 8);
  return SOSGPU_OK;
}

// SOS_GLITTER (SOS_GLITTER.F:229-371): surface-file records of a rough sea, one fused kernel (glitter_kernel.cu)
// gmodel 0: Cox-Munk glitter (wind); 1 / 2: Rondeaux / Breon BPDF (SOS_SURFACE_BPDF.F:219-392 with ISURF = 4 / 5)
static int reflection_matrices(sosgpu_ctx *ctx, int gmodel, int nbmu, const double *rmu, const double *chr, int os_nb, int os_ns,
                               int os_nm, double wind, double ind_surf, float *surf, int *il_out, double coef_c = 0.0,
                               double alpha_nadal = 0.0, double beta_nadal = 0.0, int nadal_pairing = 0)
{
  if (!ctx) return SOSGPU_ERR_NO_DEVICE;
  if (!rmu || !chr || !surf || nbmu < 1 || nbmu > SOSGPU_NBMU_MAX || os_nb > SOSGPU_NB_MAX || os_ns < 2 || os_ns > 136 ||
      os_nm < os_nb + os_ns || os_nm > 336) { ctx->err = "sosgpu_glitter: bad arguments"; return SOSGPU_ERR_ARG; }
  CK(cudaSetDevice(ctx->device));
  const int N = nbmu, W = 2 * N + 1, npair = N * (N + 1) / 2;
  const size_t nsurf = (size_t)(os_nb + 1) * 9 * N * N;
  // one block from the stream-ordered pool (no cudaMalloc / cudaFree per call): [inputs + coefficients | IL | records]
  const size_t in_bytes = ((size_t)(2 * W + 4 * (os_ns + 1)) * 8 + 255) / 256 * 256, il_bytes = ((size_t)npair * 4 + 255) / 256 * 256;
  char *d_blk = nullptr;
  CK(sos_dmalloc(ctx, &d_blk, in_bytes + il_bytes + nsurf * 4));
  struct Guard { sosgpu_ctx *c; void *p; ~Guard() { sos_dfree(c, p); } } guard{ctx, d_blk};
  double *d_in = (double *)d_blk; int *d_il = (int *)(d_blk + in_bytes); float *d_surf = (float *)(d_blk + in_bytes + il_bytes);
  CK(cudaMemcpyAsync(d_in, rmu, W * 8, cudaMemcpyHostToDevice, ctx->stream));
  CK(cudaMemcpyAsync(d_in + W, chr, W * 8, cudaMemcpyHostToDevice, ctx->stream));
  CK(cudaMemsetAsync(d_surf, 0, nsurf * 4, ctx->stream));
  std::vector<double> coef;
  const int rcf = mat_fresnel_device(ctx, N, ind_surf, os_ns, d_in, coef);
  if (rcf != SOSGPU_OK) return rcf;
  GlitterParams p{};
  p.nbmu = N; p.os_nb = os_nb; p.os_ns = os_ns; p.os_nm = os_nm; p.gmodel = gmodel; p.coef_c = coef_c;
  p.ind = ind_surf; p.alpha_nadal = alpha_nadal; p.beta_nadal = beta_nadal; p.nadal_pairing = nadal_pairing;
  p.sig = (double)0.003f + (double)0.00512f * wind;            // SIG = .003 + .00512*WIND (SOS_GLITTER.F:300)
  p.coef = gmodel == 0 ? 1.0 / p.sig : 1.0;                    // (1./SIG) (:315); SOS_MAT_REFLEXION(1.D+00, ...) for the BPDF models
  p.pi = std::acos(-1.0);
  p.rmu = d_in; p.alpha = d_in + 2 * W; p.beta = p.alpha + (os_ns + 1); p.gamma = p.beta + (os_ns + 1); p.zeta = p.gamma + (os_ns + 1);
  cudaEventRecord(ctx->ev_a, ctx->stream);
  sos_launch_glitter(p, d_surf, d_il, ctx->stream);
  cudaEventRecord(ctx->ev_b, ctx->stream);
  ctx->launches += 1;
  CK(cudaMemcpyAsync(surf, d_surf, nsurf * 4, cudaMemcpyDeviceToHost, ctx->stream));
  if (il_out) CK(cudaMemcpyAsync(il_out, d_il, npair * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  CK(cudaGetLastError());
  cudaEventElapsedTime(&ctx->last_kernel_ms, ctx->ev_a, ctx->ev_b);
  return SOSGPU_OK;
}

extern "C" int sosgpu_glitter(sosgpu_ctx *ctx, int nbmu, const double *rmu, const double *chr, int os_nb, int os_ns,
                              int os_nm, double wind, double ind_surf, float *surf, int *il_out)
{
  return reflection_matrices(ctx, 0, nbmu, rmu, chr, os_nb, os_ns, os_nm, wind, ind_surf, surf, il_out);
}

extern "C" int sosgpu_surface_bpdf(sosgpu_ctx *ctx, int isurf, int nbmu, const double *rmu, const double *chr, int os_nb, int os_ns,
                                   int os_nm, double ind_surf, double coef_c, float *surf)
{
  if (ctx && isurf != 4 && isurf != 5 && isurf != 7) {
    ctx->err = "sosgpu_surface_bpdf: the Rondeaux (4), Breon (5) and Maignan (7) models; Nadal (6) has its own entry, sosgpu_surface_nadal";
    return SOSGPU_ERR_ARG;
  }
  return reflection_matrices(ctx, isurf == 4 ? 1 : (isurf == 5 ? 2 : 3), nbmu, rmu, chr, os_nb, os_ns, os_nm, 0.0, ind_surf, surf, nullptr, coef_c);
}

// SOS_SURFACE_BPDF with ISURF = 6 (SOS_SURFACE_BPDF.F:219-392): Nadal's BPDF, series generator SOS_F21SF_NADAL (:686-1069).
// pairing 0: the surface file of the reference (its SOS_MAT_REFLEXION pairs the series by position in the file, see
// glitter_kernel.cu); 1: every pair (I, J) gets its own series.  il_out (may be NULL): series length per pair [N(N+1)/2].
extern "C" int sosgpu_surface_nadal(sosgpu_ctx *ctx, int nbmu, const double *rmu, const double *chr, int os_nb, int os_ns, int os_nm,
                                    double ind_surf, double alpha, double beta, int pairing, float *surf, int *il_out)
{
  if (ctx && (!(alpha > 0.0) || !(beta > 0.0) || os_nb < 0 || (pairing != 0 && pairing != 1))) {
    // alpha = 0 or beta = 0 make F = 0 and the reference's recombination test 0/0
    ctx->err = "sosgpu_surface_nadal: alpha and beta must be positive, pairing 0 or 1";
    return SOSGPU_ERR_ARG;
  }
  return reflection_matrices(ctx, 4, nbmu, rmu, chr, os_nb, os_ns, os_nm, 0.0, ind_surf, surf, il_out, 0.0, alpha, beta, pairing);
}

// SOS_ROUJEAN (SOS_ROUJEAN.F:212): Fourier series of Roujean's BRDF in the surface-file record layout
extern "C" int sosgpu_roujean(sosgpu_ctx *ctx, int nbmu, const double *rmu, int os_nb, double k0, double k1, double k2, float *surf)
{
  if (!ctx) return SOSGPU_ERR_NO_DEVICE;
  if (!rmu || !surf || nbmu < 1 || nbmu > SOSGPU_NBMU_MAX || os_nb < 0 || os_nb > SOSGPU_NB_MAX) { ctx->err = "sosgpu_roujean: bad arguments"; return SOSGPU_ERR_ARG; }
  CK(cudaSetDevice(ctx->device));
  const int N = nbmu, W = 2 * N + 1;
  const size_t nsurf = (size_t)(os_nb + 1) * 9 * N * N;
  const size_t in_bytes = ((size_t)W * 8 + 255) / 256 * 256, st_bytes = ((size_t)N * N * 4 + 255) / 256 * 256;
  char *d_blk = nullptr;
  CK(sos_dmalloc(ctx, &d_blk, in_bytes + st_bytes + nsurf * 4));
  struct Guard { sosgpu_ctx *c; void *p; ~Guard() { sos_dfree(c, p); } } guard{ctx, d_blk};
  double *d_rmu = (double *)d_blk; int *d_status = (int *)(d_blk + in_bytes); float *d_surf = (float *)(d_blk + in_bytes + st_bytes);
  CK(cudaMemcpyAsync(d_rmu, rmu, W * 8, cudaMemcpyHostToDevice, ctx->stream));
  CK(cudaMemsetAsync(d_surf, 0, nsurf * 4, ctx->stream));
  cudaEventRecord(ctx->ev_a, ctx->stream);
  sos_launch_roujean(N, d_rmu, os_nb, k0, k1, k2, d_surf, d_status, ctx->stream);
  cudaEventRecord(ctx->ev_b, ctx->stream);
  ctx->launches += 1;
  std::vector<int> status((size_t)N * N);
  CK(cudaMemcpyAsync(status.data(), d_status, status.size() * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaMemcpyAsync(surf, d_surf, nsurf * 4, cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  CK(cudaGetLastError());
  cudaEventElapsedTime(&ctx->last_kernel_ms, ctx->ev_a, ctx->ev_b);
  for (int v : status)
    if (v) { ctx->err = "SOS_FSF_ROUJEAN : BRDF < 0 (unsuitable model parameters)"; return SOSGPU_ERR_IER; }   // label 993
  return SOSGPU_OK;
}

// SOS_BPDF_AJOUT_BRDF (SOS_SURFACE.F:2503): out = surf1 + surf2 over [os_nb+1][9][N][N] REAL*4 records
extern "C" int sosgpu_bpdf_ajout_brdf(sosgpu_ctx *ctx, const float *surf1, const float *surf2, int nbmu, int os_nb, float *out)
{
  if (!ctx) return SOSGPU_ERR_NO_DEVICE;
  if (!surf1 || !surf2 || !out || nbmu < 1 || nbmu > SOSGPU_NBMU_MAX || os_nb < 0 || os_nb > SOSGPU_NB_MAX) return SOSGPU_ERR_ARG;
  CK(cudaSetDevice(ctx->device));
  const size_t n = (size_t)(os_nb + 1) * 9 * nbmu * nbmu;
  float *d = nullptr;
  CK(sos_dmalloc(ctx, &d, 3 * n * 4));
  struct Guard { sosgpu_ctx *c; void *p; ~Guard() { sos_dfree(c, p); } } guard{ctx, d};
  CK(cudaMemcpyAsync(d, surf1, n * 4, cudaMemcpyHostToDevice, ctx->stream));
  CK(cudaMemcpyAsync(d + n, surf2, n * 4, cudaMemcpyHostToDevice, ctx->stream));
  sos_launch_ajout_brdf(d + 2 * n, d, d + n, n, ctx->stream);
  ctx->launches += 1;
  CK(cudaMemcpyAsync(out, d + 2 * n, n * 4, cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  CK(cudaGetLastError());
  return SOSGPU_OK;
}

// ---------------------------------------------------------------------------------------------
// gfortran-ABI drop-ins.  One process-wide context, created on first use.
// (device: SOSGPU_DEVICE, default 0); shared by the shims of sosgpu_profile.cu and sosgpu_aerosols.cu.
static sosgpu_ctx *g_ctx = nullptr;
sosgpu_ctx *sos_shim_ctx()
{
  if (!g_ctx) {
    const char *e = getenv("SOSGPU_DEVICE");
    if (sosgpu_create(&g_ctx, e ? atoi(e) : 0) != SOSGPU_OK) {
      fprintf(stderr, "  libsosgpu: no usable CUDA device -- the SOS hot path has no CPU fallback\n");
      g_ctx = nullptr;
    }
  }
  return g_ctx;
}
static sosgpu_ctx *shim_ctx() { return sos_shim_ctx(); }

static std::string fstr(const char *s, size_t n)
{
  std::string r(s, n);
  const size_t sp = r.find(' ');                               // the reference cuts file names at the first blank
  if (sp != std::string::npos) r.resize(sp);
  return r;
}

// gfortran unformatted sequential: 4-byte record markers
static bool read_records(const std::string &path, std::vector<std::vector<char>> &recs)
{
  FILE *f = fopen(path.c_str(), "rb");
  if (!f) return false;
  for (;;) {
    int32_t n = 0;
    if (fread(&n, 4, 1, f) != 1) break;
    std::vector<char> r((size_t)n);
    if (n > 0 && fread(r.data(), 1, (size_t)n, f) != (size_t)n) { fclose(f); return false; }
    int32_t m = 0;
    if (fread(&m, 4, 1, f) != 1 || m != n) { fclose(f); return false; }
    recs.push_back(std::move(r));
  }
  fclose(f);
  return true;
}
static bool write_records(const std::string &path, const double *rec, int nrec, size_t doubles_per_rec)
{
  FILE *f = fopen(path.c_str(), "wb");
  if (!f) return false;
  const int32_t n = (int32_t)(doubles_per_rec * 8);
  for (int r = 0; r < nrec; ++r) {
    fwrite(&n, 4, 1, f);
    fwrite(rec + (size_t)r * doubles_per_rec, 8, doubles_per_rec, f);
    fwrite(&n, 4, 1, f);
  }
  fclose(f);
  return true;
}

#define NBMU_MAX SOSGPU_NBMU_MAX
#define NB_MAX SOSGPU_NB_MAX
#define NT_MAX SOSGPU_NT_MAX
#define SOSGPU_THRESHOLD_Q_U_NULL 1.0e-15   /* CTE_THRESHOLD_Q_U_NULL, SOS.h:418 */

extern "C" int sosgpu_batch_upload_os(sosgpu_ctx *ctx, const sosgpu_optics *optics, int noptics,
                                      const sosgpu_term *terms, int nterm, int ngroup, const int *iborm,
                                      sosgpu_batch **batch);

// SOS_OS.F:303-308.  Arrays use the reference's fixed extents: RMU/GA(-80:80), H/XDEL/YDEL/ZPROF(0:600),
// ALPHA..ZETA(0:200).  TRACE/IDLOG are accepted and ignored (no log unit on this side).
extern "C" void sos_os_(const int *nbmu, double *rmu, const double *ga, const int *os_nb, const int *nt,
                        const char *ficsurf, const char *ficos,
                        const int *n0, const double *tetas, const double *ro, const int *imat_surf,
                        const int *ifresnel, const double *ind_surf,
                        const double *h, const double *xdel, const double *ydel, const double *zprof, const double *ron,
                        double *alpha, double *beta, double *gamma, double *zeta, const double *zout,
                        const int *igmax, const int *iborm, const int *ipolar, const int *trace, const int *idlog,
                        double *emoins, double *eplus, int *ier, size_t len_ficsurf, size_t len_ficos)
{
  (void)trace; (void)idlog; (void)beta;
  *ier = 0;
  const int N = *nbmu, W = 2 * N + 1, NB = *os_nb, NT = *nt;
  sosgpu_ctx *ctx = shim_ctx();
  if (!ctx || N < 1 || N > NBMU_MAX || NB > NB_MAX || NT > NT_MAX) { *ier = -1; return; }
  const std::string fsurf = fstr(ficsurf, len_ficsurf), fos = fstr(ficos, len_ficos);
  std::vector<double> rmu_c(W), ga_c(W);
  for (int j = -N; j <= N; ++j) { rmu_c[j + N] = rmu[j + NBMU_MAX]; ga_c[j + N] = ga[j + NBMU_MAX]; }
  std::vector<float> surf;
  int nsurf = 0;
  if (*imat_surf == 1) {                                       // :664-666, :916-925
    std::vector<std::vector<char>> recs;
    if (!read_records(fsurf, recs)) { fprintf(stdout, "  ERROR on SURFACE file opening for SOS_OS\n"); *ier = -1; return; }
    const size_t need = (size_t)9 * N * N * 4;
    for (auto &r : recs) {
      if (r.size() < need) { fprintf(stdout, "  ERROR on SURFACE file reading for SOS_OS\n"); *ier = -1; return; }
      const float *p = (const float *)r.data();
      surf.insert(surf.end(), p, p + (size_t)9 * N * N);
      ++nsurf;
    }
    if (nsurf < *iborm + 1) { fprintf(stdout, "  ERROR on SURFACE file reading for SOS_OS\n"); *ier = -1; return; }
  }
  sosgpu_optics o{};
  o.nbmu = N; o.rmu = rmu_c.data(); o.ga = ga_c.data(); o.n0 = *n0; o.tetas = *tetas; o.os_nb = NB;
  o.alpha = alpha; o.beta = beta; o.gamma = gamma; o.zeta = zeta; o.a_trunc = 0.0; o.piz = 1.0; o.piztr = 1.0;
  o.ron = *ron; o.rho = *ro; o.imat_surf = *imat_surf; o.ifresnel = *ifresnel; o.ind_surf = *ind_surf;
  o.surf = surf.empty() ? nullptr : surf.data(); o.n_surf_rec = nsurf; o.igmax = *igmax; o.ipolar = *ipolar; o.zout = *zout;
  sosgpu_term t{};
  t.optics = 0; t.group = 0; t.aik = 1.0; t.nt = NT; t.zprof = zprof; t.h = h; t.pcaer = xdel; t.pcmol = ydel;
  // caller-visible side effects (:693-697, :715)
  if (*ipolar == 0) for (int k = 0; k <= NB; ++k) { alpha[k] = 0.0; gamma[k] = 0.0; zeta[k] = 0.0; }
  const double tab = (*n0 > 0) ? -rmu_c[*n0 + N] : -std::cos(std::acos(-1.0) * (*tetas) / 180.0);
  rmu[NBMU_MAX] = tab;
  if (tab == 0.0) { fprintf(stdout, "  SOS computation are stopped   because of a limb incidence\n"); return; }
  if ((((*zout) < 0) && ((*zout) != -1.0)) || ((*zout) > 120.0)) {
    fprintf(stdout, "  Inconsistency output altitude in the   atmospheric profile\n");
    *ier = -1; return;
  }
  sosgpu_batch *b = nullptr;
  const int ib = *iborm;
  int rc = sosgpu_batch_upload_os(ctx, &o, 1, &t, 1, 1, &ib, &b);
  if (rc != SOSGPU_OK) { *ier = -1; return; }
  const int rs = NB + 1;
  std::vector<double> rec((size_t)rs * 3 * W, 0.0);
  int nf = 0; double em = 0, ep = 0;
  sosgpu_term_out to{};
  to.rec = rec.data(); to.n_fourier = &nf; to.emoins = &em; to.eplus = &ep;
  rc = sosgpu_batch_run(ctx, b, rs, W, 0, &to, nullptr);
  sosgpu_batch_free(ctx, b);
  if (rc != SOSGPU_OK) { fprintf(stderr, "  libsosgpu: %s\n", sosgpu_last_error(ctx)); *ier = -1; return; }
  *emoins = em; *eplus = ep;
  if (fos != "NO_OUTPUT") {                                    // :671, :1571-1575
    if (!write_records(fos, rec.data(), nf, (size_t)3 * W)) {
      fprintf(stdout, "  ERROR on SOS binary result file opening for SOS_OS\n");
      *ier = -1;
    }
  }
}

// SOS_AGGREGATE.F:172-178: RES = RES + AIK*TMP record by record with zero padding, file read-modify-write,
// then the scalar accumulators (:452-488).  TDIFMUG uses the (-80:80) extent.
extern "C" void sos_aggregate_(const int *nbmu, const double *aik, const char *ficos_tmp,
                               const double *ttot_tronc_tmp, const double *ttot_vrai_tmp, const double *tauout_tmp,
                               const double *tdifmus_tmp, const double *tdifmug_tmp, const double *emoins_tmp,
                               const double *eplus_tmp, const char *ficos_agg_tmp, const char *ficos,
                               double *ttot_tronc, double *ttot_vrai, double *tauout, double *tdifmus, double *tdifmug,
                               double *emoins, double *eplus, int *ier,
                               size_t len_ficos_tmp, size_t len_ficos_agg_tmp, size_t len_ficos)
{
  (void)ficos_agg_tmp; (void)len_ficos_agg_tmp;
  const int N = *nbmu, W = 2 * N + 1;
  sosgpu_ctx *ctx = shim_ctx();
  if (!ctx) { *ier = -1; return; }
  const std::string ftmp = fstr(ficos_tmp, len_ficos_tmp), fres = fstr(ficos, len_ficos);
  std::vector<std::vector<char>> rt, rr;
  if (!read_records(ftmp, rt)) { fprintf(stdout, "  SOS_AGGREGATE : ERROR_1101 : \n"); *ier = -1; return; }
  const bool have = read_records(fres, rr);
  const size_t per = (size_t)3 * W;
  // record count incl. the reference's trailing all-zero record when the existing file is not shorter (:372-393)
  size_t nout = std::max(rt.size(), rr.size());
  if (have && rr.size() >= rt.size()) nout += 1;
  std::vector<double> a(nout * per, 0.0), r(nout * per, 0.0);
  for (size_t i = 0; i < rt.size(); ++i) memcpy(&a[i * per], rt[i].data(), std::min(rt[i].size(), per * 8));
  for (size_t i = 0; i < rr.size(); ++i) memcpy(&r[i * per], rr[i].data(), std::min(rr[i].size(), per * 8));
  double *d_a = nullptr, *d_r = nullptr;
  bool ok = cudaSetDevice(ctx->device) == cudaSuccess && cudaMalloc(&d_a, a.size() * 8) == cudaSuccess &&
            cudaMalloc(&d_r, r.size() * 8) == cudaSuccess;
  if (ok) {
    cudaMemcpyAsync(d_a, a.data(), a.size() * 8, cudaMemcpyHostToDevice, ctx->stream);
    cudaMemcpyAsync(d_r, r.data(), r.size() * 8, cudaMemcpyHostToDevice, ctx->stream);
    sos_launch_axpy(d_r, d_a, *aik, r.size(), ctx->stream);
    ctx->launches += 1;
    cudaMemcpyAsync(r.data(), d_r, r.size() * 8, cudaMemcpyDeviceToHost, ctx->stream);
    ok = cudaStreamSynchronize(ctx->stream) == cudaSuccess && cudaGetLastError() == cudaSuccess;
  }
  cudaFree(d_a); cudaFree(d_r);
  if (!ok || !write_records(fres, r.data(), (int)nout, per)) { fprintf(stdout, "  SOS_AGGREGATE : ERROR_1302 : \n"); *ier = -1; return; }
  *tdifmus = *tdifmus + (*aik) * (*tdifmus_tmp);               // :452-459
  *emoins = *emoins + (*aik) * (*emoins_tmp);
  *eplus = *eplus + (*aik) * (*eplus_tmp);
  for (int j = -N; j <= N; ++j) tdifmug[j + NBMU_MAX] = tdifmug[j + NBMU_MAX] + (*aik) * tdifmug_tmp[j + NBMU_MAX];
  double trans;                                                // :467-488
  trans = (*ttot_tronc != 0) ? (*aik) * std::exp(-*ttot_tronc_tmp) + std::exp(-*ttot_tronc) : (*aik) * std::exp(-*ttot_tronc_tmp);
  *ttot_tronc = -std::log(trans);
  trans = (*ttot_vrai != 0) ? (*aik) * std::exp(-*ttot_vrai_tmp) + std::exp(-*ttot_vrai) : (*aik) * std::exp(-*ttot_vrai_tmp);
  *ttot_vrai = -std::log(trans);
  trans = (*tauout != 0) ? (*aik) * std::exp(-*tauout_tmp) + std::exp(-*tauout) : (*aik) * std::exp(-*tauout_tmp);
  *tauout = -std::log(trans);
}

// ---------------------------------------------------------------------------------------------
// SOS.F:340-345.  Reads PROFIL_TMP (format 70: 2X,I5,F10.5,3(E15.8), :511-516), applies the truncation
// adaptation (:523-543), solves through sos_os_ (so FICOS / FICSURF and the caller-visible side effects behave
// as there), derives TAUOUT / TTOT_* (:518,567-586) and, when FICTRANS is not 'NO_OUTPUT', the 1+N black-surface
// IBORM=0 solves of the -SOS.Trans branch (:605-637) as ONE device batch.
static bool parse_field(const std::string &line, size_t pos, size_t len, double &v)
{
  if (line.size() < pos + 1) return false;
  std::string f = line.substr(pos, len);
  for (auto &c : f) if (c == 'D' || c == 'd') c = 'E';
  char *end = nullptr;
  v = std::strtod(f.c_str(), &end);
  return end != f.c_str();
}

extern "C" void sos_(const char *ficos, const char *fictrans, const char *ficprofil,
                     const int *nt, const double *zout, const int *igmax, const int *ipolar, const double *ron,
                     const double *ind_surf, const double *rho, const int *imat_surf, const int *ifresnel,
                     const char *ficsurf, const int *n0, const double *piz, const double *piztr, const double *a,
                     double *rmu, const double *ga, const double *tetas, const int *os_nb, const int *lum_nbmu,
                     double *alpha, double *beta, double *gamma, double *zeta,
                     double *ttot_tronc, double *ttot_vrai, double *tauout, double *tdifmus, double *tdifmug,
                     double *emoins, double *eplus, const int *trace, const int *idlog, int *ier,
                     size_t len_ficos, size_t len_fictrans, size_t len_ficprofil, size_t len_ficsurf)
{
  *ier = 0;
  const int NT = *nt, N = *lum_nbmu, NB = *os_nb;
  if (NT < 1 || NT > NT_MAX || N < 1 || N > NBMU_MAX || NB > NB_MAX) { *ier = -1; return; }
  std::vector<double> zprof(NT_MAX + 1, 0.0), h(NT_MAX + 1, 0.0), xdel(NT_MAX + 1, 0.0), ydel(NT_MAX + 1, 0.0),
      htr(NT_MAX + 1, 0.0);
  {
    FILE *f = fopen(fstr(ficprofil, len_ficprofil).c_str(), "r");
    if (!f) { fprintf(stdout, "  ERROR on PROFILE file opening for SOS\n"); *ier = -1; return; }
    char buf[512];
    for (int i = 0; i <= NT; ++i) {
      bool ok = fgets(buf, sizeof buf, f) != nullptr;
      if (ok) {
        const std::string line(buf);
        ok = parse_field(line, 7, 10, zprof[i]) && parse_field(line, 17, 15, h[i]) &&
             parse_field(line, 32, 15, xdel[i]) && parse_field(line, 47, 15, ydel[i]);
      }
      if (!ok) { fclose(f); fprintf(stdout, "  ERROR on PROFILE file reading for SOS\n"); *ier = -1; return; }
    }
    fclose(f);
  }
  *ttot_vrai = h[NT];                                          // :518
  bool lta = true;
  htr[0] = h[0];
  if (*a != 0.0) {                                             // :525-537
    for (int i = 1; i <= NT; ++i) {
      const double dh = h[i] - h[i - 1];
      const double va = xdel[i] * dh;
      const double vatr = va * (1 - (*piz) * 0.5 * (*a));
      const double vr = ydel[i] * dh;
      const double vg = (1 - xdel[i] - ydel[i]) * dh;
      htr[i] = (vatr + vr + vg) + htr[i - 1];
      xdel[i] = vatr / (vatr + vr + vg);
      ydel[i] = vr / (vatr + vr + vg);
    }
  }
  for (int i = 0; i <= NT; ++i) {                              // :539-543
    if (*a != 0.0) h[i] = htr[i];
    xdel[i] = xdel[i] * (*piztr);
    if (xdel[i] != 0.0) lta = false;
  }
  const int iborm = lta ? 2 : NB;                              // :549-550
  sos_os_(lum_nbmu, rmu, ga, os_nb, nt, ficsurf, ficos, n0, tetas, rho, imat_surf, ifresnel, ind_surf,
          h.data(), xdel.data(), ydel.data(), zprof.data(), ron, alpha, beta, gamma, zeta, zout, igmax, &iborm, ipolar,
          trace, idlog, emoins, eplus, ier, len_ficsurf, len_ficos);
  if (*ier != 0) { fprintf(stdout, "  ERROR on subroutine SOS_OS\n"); *ier = -1; return; }
  if (*zout == -1) *tauout = h[0];                             // :567-583
  else {
    int j = 1;
    while (j < NT && *zout < zprof[j]) ++j;
    const double zz = (*zout - zprof[j - 1]) / (zprof[j] - zprof[j - 1]);
    *tauout = (1 - zz) * h[j - 1] + zz * h[j];
  }
  *ttot_tronc = h[NT];                                         // :586
  if (fstr(fictrans, len_fictrans) == "NO_OUTPUT") return;

  // -SOS.Trans (:605-637): SOS_OS with RO=0, no surface matrix, no Fresnel, ZOUT=-1, IBORM=0 for the solar
  // direction and then for N0 = 1..N; TDIFMUS / TDIFMUG(J) are the EMOINS of those solves.
  sosgpu_ctx *ctx = shim_ctx();
  if (!ctx) { *ier = -1; return; }
  const int W = 2 * N + 1;
  std::vector<double> rmu_c(W), ga_c(W);
  for (int j = -N; j <= N; ++j) { rmu_c[j + N] = rmu[j + NBMU_MAX]; ga_c[j + N] = ga[j + NBMU_MAX]; }
  std::vector<sosgpu_optics> ov(N + 1);
  std::vector<sosgpu_term> tv(N + 1);
  std::vector<int> ib(N + 1, 0);
  for (int j = 0; j <= N; ++j) {
    sosgpu_optics &o = ov[j];
    o = sosgpu_optics{};
    o.nbmu = N; o.rmu = rmu_c.data(); o.ga = ga_c.data(); o.n0 = (j == 0) ? *n0 : j; o.tetas = *tetas; o.os_nb = NB;
    o.alpha = alpha; o.beta = beta; o.gamma = gamma; o.zeta = zeta; o.a_trunc = 0.0; o.piz = 1.0; o.piztr = 1.0;
    o.ron = *ron; o.rho = 0.0; o.imat_surf = 0; o.ifresnel = 0; o.ind_surf = *ind_surf; o.surf = nullptr; o.n_surf_rec = 0;
    o.igmax = *igmax; o.ipolar = *ipolar; o.zout = -1.0;
    sosgpu_term &t = tv[j];
    t = sosgpu_term{};
    t.optics = j; t.group = 0; t.aik = 1.0; t.nt = NT; t.zprof = zprof.data(); t.h = h.data(); t.pcaer = xdel.data();
    t.pcmol = ydel.data();
  }
  sosgpu_batch *b = nullptr;
  int rc = sosgpu_batch_upload_os(ctx, ov.data(), N + 1, tv.data(), N + 1, 1, ib.data(), &b);
  std::vector<double> em(N + 1, 0.0);
  std::vector<int> ierv(N + 1, 0);
  if (rc == SOSGPU_OK) {
    sosgpu_term_out to{};
    to.emoins = em.data(); to.ier = ierv.data();
    rc = sosgpu_batch_run(ctx, b, NB + 1, W, 0, &to, nullptr);
    sosgpu_batch_free(ctx, b);
  }
  for (int j = 0; j <= N && rc == SOSGPU_OK; ++j) if (ierv[j] != 0) rc = SOSGPU_ERR_IER;
  if (rc != SOSGPU_OK) { fprintf(stdout, "  ERROR on subroutine SOS_OS\n"); *ier = -1; return; }
  *tdifmus = em[0];
  for (int j = 1; j <= N; ++j) tdifmug[j + NBMU_MAX] = em[j];
  rmu[NBMU_MAX] = -rmu[N + NBMU_MAX];                          // RMU(0) after the last call (N0 = LUM_NBMU, SOS_OS.F:715)
}

// SOS_GLITTER.F:229-233.  The three intermediate files of the reference (RES_GSF, RES_FRESNEL, RES_MAT_REFLEX) are
// deleted by the reference before it returns (:363-368) and are never created here; FICGLITTER must not pre-exist
// (STATUS='NEW', SOS_SURFACE.F:2360) and receives OS_NB+1 unformatted records of 9*N*N REAL*4 (:2404-2412).
extern "C" void sos_glitter_(const int *lum_nbmu, const double *rmu, const double *chr, const double *wind,
                             const double *ind, const int *os_nb, const int *os_ns, const int *os_nm,
                             const char *fic_res_gsf, const char *fic_res_fresnel, const char *fic_res_mat_reflex,
                             const char *ficglitter, const int *trace, int *ier,
                             size_t len_gsf, size_t len_fresnel, size_t len_mat, size_t len_ficglitter)
{
  (void)fic_res_gsf; (void)fic_res_fresnel; (void)fic_res_mat_reflex; (void)len_gsf; (void)len_fresnel; (void)len_mat;
  (void)trace;
  *ier = 0;
  const int N = *lum_nbmu, W = 2 * N + 1, NB = *os_nb;
  sosgpu_ctx *ctx = shim_ctx();
  if (!ctx || N < 1 || N > NBMU_MAX || NB > NB_MAX) { *ier = -1; return; }
  const std::string fout = fstr(ficglitter, len_ficglitter);
  if (FILE *f = fopen(fout.c_str(), "rb")) {
    fclose(f);
    fprintf(stdout, "  ERROR on a file opening for SOS_MISE_FORMAT\n Erreur dans la routine SOS_MISE_FORMAT\n");
    *ier = -1; return;
  }
  std::vector<double> rmu_c(W), chr_c(W);
  for (int j = -N; j <= N; ++j) { rmu_c[j + N] = rmu[j + NBMU_MAX]; chr_c[j + N] = chr[j + NBMU_MAX]; }
  const size_t per = (size_t)9 * N * N;
  std::vector<float> surf(per * (NB + 1));
  const int rc = sosgpu_glitter(ctx, N, rmu_c.data(), chr_c.data(), NB, *os_ns, *os_nm, *wind, *ind, surf.data(), nullptr);
  if (rc != SOSGPU_OK) { fprintf(stdout, " Erreur dans la routine SOS_MAT_REFLEXION\n"); *ier = -1; return; }
  FILE *f = fopen(fout.c_str(), "wb");
  if (!f) { fprintf(stdout, "  ERROR on a file opening for SOS_MISE_FORMAT\n"); *ier = -1; return; }
  const int32_t n = (int32_t)(per * 4);
  for (int s = 0; s <= NB; ++s) {
    fwrite(&n, 4, 1, f);
    fwrite(&surf[(size_t)s * per], 4, per, f);
    fwrite(&n, 4, 1, f);
  }
  fclose(f);
}

// Reads FICHOS (records Q,U,I(-N:N), SOS_TRPHI.F:895-897) into the compact layout
static bool read_fichos(const std::string &path, int N, std::vector<double> &rec, int &nrec)
{
  std::vector<std::vector<char>> recs;
  if (!read_records(path, recs)) { fprintf(stdout, "SOS_TRPHI : Error while opening a file\n"); return false; }
  const size_t per = (size_t)3 * (2 * N + 1);
  nrec = (int)recs.size();
  rec.assign(per * std::max(nrec, 1), 0.0);
  for (int i = 0; i < nrec; ++i) {
    if (recs[i].size() < per * 8) { fprintf(stdout, "SOS_TRPHI : Error while reading or writing on a file\n"); return false; }
    memcpy(&rec[i * per], recs[i].data(), per * 8);
  }
  if (nrec < 1) { fprintf(stdout, "SOS_TRPHI : Error while reading or writing on a file\n"); return false; }
  return true;
}

// the land-surface direct terms requested through the Fortran argument list, for the duration of one shim call
struct DirectModelScope {
  sosgpu_ctx *ctx; sosgpu_direct_models saved;
  DirectModelScope(sosgpu_ctx *c, const int *iroujean, const double *k0, const double *k1, const double *k2, const int *irondeaux,
                   const int *ibreon, const int *inadal, const double *alpha_nadal, const double *beta_nadal, const int *imaignan,
                   const double *coef_c_maignan) : ctx(c), saved(c->dm)
  {
    ctx->dm = sosgpu_direct_models{*iroujean, *k0, *k1, *k2, *irondeaux, *ibreon, *inadal, *alpha_nadal, *beta_nadal, *imaignan,
                                   *coef_c_maignan};
  }
  ~DirectModelScope() { ctx->dm = saved; }
};

// SOS_TRPHI.F:749-755: one azimuth PHI (radians).  XIT/XQT/XUT/ANGDIFF(-80:80); index 0 of the Stokes arrays only
// sees the final thresholding (:1212-1218), as in the reference.
extern "C" void sos_trphi_(const char *fichos, const int *nbmu, const double *rmu, const double *tau,
                           const double *tauout, const double *phi, const int *igli, const int *n0, const double *wind,
                           const double *ind_surf, const int *ifresnel, const int *iroujean, const double *k0,
                           const double *k1, const double *k2, const int *irondeaux, const int *ibreon,
                           const int *inadal, const double *alpha_nadal, const double *beta_nadal, const int *imaignan,
                           const double *coef_c_maignan, const int *ipolar, double *xit, double *xqt, double *xut,
                           double *angdiff, int *ier, size_t len_fichos)
{
  *ier = 0;
  const int N = *nbmu, W = 2 * N + 1;
  sosgpu_ctx *ctx = shim_ctx();
  if (!ctx || N < 1 || N > NBMU_MAX || *n0 < 1 || *n0 > N) { *ier = -1; return; }
  DirectModelScope dms(ctx, iroujean, k0, k1, k2, irondeaux, ibreon, inadal, alpha_nadal, beta_nadal, imaignan, coef_c_maignan);
  std::vector<double> rec, rmu_c(W), out;
  int nrec = 0;
  if (!read_fichos(fstr(fichos, len_fichos), N, rec, nrec)) { *ier = -1; return; }
  for (int j = -N; j <= N; ++j) rmu_c[j + N] = rmu[j + NBMU_MAX];
  const std::vector<double> phis(1, *phi);
  if (trphi_core(ctx, rec.data(), nrec, N, rmu_c.data(), *tau, *tauout, *igli, *n0, *wind, *ind_surf, *ifresnel, *ipolar,
                 phis, out) != SOSGPU_OK) { *ier = -1; return; }
  for (int j = 1; j <= N; ++j) {
    angdiff[NBMU_MAX + j] = out[(size_t)(0 * 7 + 0) * N + j - 1]; angdiff[NBMU_MAX - j] = out[(size_t)(1 * 7 + 0) * N + j - 1];
    xit[NBMU_MAX + j] = out[(size_t)(0 * 7 + 1) * N + j - 1];     xit[NBMU_MAX - j] = out[(size_t)(1 * 7 + 1) * N + j - 1];
    xqt[NBMU_MAX + j] = out[(size_t)(0 * 7 + 2) * N + j - 1];     xqt[NBMU_MAX - j] = out[(size_t)(1 * 7 + 2) * N + j - 1];
    xut[NBMU_MAX + j] = out[(size_t)(0 * 7 + 3) * N + j - 1];     xut[NBMU_MAX - j] = out[(size_t)(1 * 7 + 3) * N + j - 1];
  }
  const double pi = std::acos(-1.0), c0 = rmu[NBMU_MAX + *n0];   // :882-891 for J = 0
  angdiff[NBMU_MAX] = std::acos(-c0 * rmu[NBMU_MAX] + std::sin(std::acos(c0)) * std::sin(std::acos(rmu[NBMU_MAX])) * std::cos(*phi))
                      * 180.0 / pi;
  if (xit[NBMU_MAX] <= 1e-99) xit[NBMU_MAX] = 0.0;
  if (std::fabs(xqt[NBMU_MAX]) < SOSGPU_THRESHOLD_Q_U_NULL) xqt[NBMU_MAX] = 0.0;
  if (std::fabs(xut[NBMU_MAX]) < SOSGPU_THRESHOLD_Q_U_NULL) xut[NBMU_MAX] = 0.0;
}

// SOS_TRPHI.F:285-300.  Output tables use the reference's extents: PHI_FIN(0:360), THETA_FIN(0:80), X_FIN(0:360,0:80)
// (element (IP,JJ) at IP + 361*JJ).
extern "C" void sos_trphi_option_(const int *nbmu, const double *rmu, const double *ga, const char *fichos,
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
                                  double *pol_rate_down, double *l_pol_down, int *ier, size_t len_fichos)
{
  (void)ga; (void)zout;
  *ier = 0;
  const int N = *nbmu, W = 2 * N + 1;
  sosgpu_ctx *ctx = shim_ctx();
  if (!ctx || N < 1 || N > NBMU_MAX) { *ier = -1; return; }
  DirectModelScope dms(ctx, iroujean, k0, k1, k2, irondeaux, ibreon, inadal, alpha_nadal, beta_nadal, imaignan, coef_c_maignan);
  if (*itrphi != 1 && *itrphi != 2) return;                      // neither branch runs in the reference (:431,556)
  std::vector<double> rec, rmu_c(W);
  int nrec = 0;
  if (!read_fichos(fstr(fichos, len_fichos), N, rec, nrec)) { *ier = -1; return; }
  for (int j = -N; j <= N; ++j) rmu_c[j + N] = rmu[j + NBMU_MAX];
  const int cap = 361;
  std::vector<double> pf(cap), tf(N), up((size_t)7 * cap * N), down((size_t)7 * cap * N);
  const int nphi = sosgpu_trphi_option(ctx, rec.data(), nrec, N, rmu_c.data(), *tau, *tauout, *igli, *n0, *wind, *ind_surf,
                                       *ifresnel, *itrphi, *phios, *pas_phi, *ipolar, pf.data(), tf.data(), up.data(),
                                       down.data(), cap);
  if (nphi < 1) { *ier = -1; return; }
  double *dst_up[7] = {sca_up, i_up, q_up, u_up, pol_ang_up, pol_rate_up, l_pol_up};
  double *dst_dn[7] = {sca_down, i_down, q_down, u_down, pol_ang_down, pol_rate_down, l_pol_down};
  for (int t = 0; t < 7; ++t)
    for (int ip = 0; ip < nphi; ++ip)
      for (int jj = 0; jj < N; ++jj) {
        dst_up[t][ip + (size_t)361 * jj] = up[((size_t)t * cap + ip) * N + jj];
        dst_dn[t][ip + (size_t)361 * jj] = down[((size_t)t * cap + ip) * N + jj];
      }
  if (*itrphi == 1) phi_fin[0] = pf[0];                          // only PHI_FIN(0) is assigned for ITRPHI=1 (:445,504)
  else for (int ip = 0; ip < nphi; ++ip) phi_fin[ip] = pf[ip];
  for (int jj = 0; jj < N; ++jj) theta_fin[jj] = tf[jj];
}
