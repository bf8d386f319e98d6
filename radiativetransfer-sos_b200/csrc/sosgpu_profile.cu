// sosgpu_profile.cu -- the per-term profile chain of libsosgpu.so: for every (wavelength, CKD term) of a
// band, the gas absorption profile (SOS_ABSPROFILE + COEFF_ABS_CKD), the discretisation of the atmosphere in optical depth
// (SOS_PROFILE + SOS_DISC) and the decimal round trip of the PROFIL_TMP text file, as three kernels over the whole band.  The
// reference runs this chain serially on the host before each term-solve (SOS_PROC.F:3494-3537) and hands the result to SOS
// through a text file; here its outputs are the arrays sosgpu_term points to.  Also: the CKD table reader READ_CKD_COEFF
// (host: it parses text files) and the gfortran-ABI symbols of the three reference routines.
// Compiled with -fmad=false: NT is an integer result of threshold comparisons, the arithmetic rounds as the reference's.
#include "sosgpu_host.h"
#include "profile_chain.cuh"
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <algorithm>
#include <map>

namespace {

struct ProfTermDev {          // one (wavelength, CKD term) on the device
  int lamb;                   // 1-based index into the PACKED spectral slices
  int ik[PC_NBABS];
  int absprofil, iprofil;
  double tr, hr, ta, ha, zmin, zmax;
};

// SOS_ABSPROFILE: one block per term, one thread per gas layer (the 8 gases of a layer are summed in the reference's order by
// one thread), then the running transmission by thread 0.  tauabs [nterm][50].
__global__ void __launch_bounds__(64)
k_absprofile(PcCkd ckd, const double *__restrict__ userprofil, const double *__restrict__ ro, const ProfTermDev *__restrict__ terms,
             double *__restrict__ tauabs, int *__restrict__ ier)
{
  __shared__ double tau_layer[PC_NLEV];
  __shared__ int err;
  const ProfTermDev &t = terms[blockIdx.x];
  const int j = threadIdx.x + 1;                                   // layer 1 (top) .. 49
  if (threadIdx.x == 0) err = 0;
  __syncthreads();
  double *out = tauabs + (size_t)blockIdx.x * PC_NLEV;
  if (t.absprofil == 7) {                                          // SOS_ABSPROFILE.F:318: no absorption
    if (threadIdx.x < PC_NLEV) out[threadIdx.x] = 0.0;
    return;
  }
  if (j <= PC_NLEV - 1) {
    double tau = 0.0;
    const int rc = pc_absprofile_layer(ckd, userprofil, ro, t.lamb, t.ik, j, &tau);
    if (rc != 0) atomicMax(&err, rc);
    tau_layer[j - 1] = tau;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    if (err != 0) { ier[blockIdx.x] = err; for (int i = 0; i < PC_NLEV; ++i) out[i] = 0.0; }
    else pc_absprofile_scan(tau_layer, out);
  }
}

// The search strategy of the kernels: one warp per term.  SOS_DISC's bisection advances five levels per round (31 lanes
// evaluate the candidates of a depth-5 subtree, pc_tree_walk reads the serial path off their results) and the first-level scan
// 32 steps of CTE_DELTA_Z per round.  Every candidate is the double the serial loop would have tested (tests/profile_host.cpp
// runs this strategy lane by lane on the host against the reference), so NT and the levels do not depend on the strategy; what
// changes is the length of the dependent chain of FP64 exp / divide evaluations per profile: about 8000 -> 1200.
struct PcWarp {
  __host__ __device__ double disc(double dt, const PcColumn &c, double tim1, double zmax_init, double zlim) const
  {
#ifdef __CUDA_ARCH__
    const int lane = threadIdx.x & 31;
    const int node = lane ? lane : 1;                              // lane 0 repeats the root; bit 0 of the masks is never read
    const int depth = 31 - __clz(node);
    const double ti = tim1 + dt;
    double zmax = zmax_init, zmin = zlim;
    for (;;) {
      double lo, hi;
      pc_tree_interval(node, depth, zmin, zmax, &lo, &hi);
      const double cand = (hi + lo) / 2.0;
      const int r = pc_disc_step(c, ti, cand);
      const unsigned stop = __ballot_sync(0xffffffffu, r & 1), dir = __ballot_sync(0xffffffffu, r & 2);
      int n;
      if (pc_tree_walk(stop, dir, &n)) return __shfl_sync(0xffffffffu, cand, n);
      pc_tree_interval(n, 5, zmin, zmax, &zmin, &zmax);
    }
#else
    (void)dt; (void)c; (void)tim1; (void)zmax_init; (void)zlim;
    return 0.0;                                                    // never instantiated for the host
#endif
  }
  __host__ __device__ void first(bool gas, const PcColumn &c, double t_first, double *z, double *dtau) const
  {
#ifdef __CUDA_ARCH__
    if (!(0.0 < t_first)) return;
    const int lane = threadIdx.x & 31;
    double z0 = *z;
    for (;;) {
      double zz = z0;
      for (int k = 0; k <= lane; ++k) zz = zz - PC_DELTA_Z;        // the reference's repeated subtraction, so the same doubles
      const double d = gas ? pc_first_tau_gas(c, zz) : pc_first_tau_ng(c, zz);
      const unsigned hit = __ballot_sync(0xffffffffu, !(d < t_first));
      if (hit) {
        const int l = __ffs(hit) - 1;
        *z = __shfl_sync(0xffffffffu, zz, l);
        *dtau = __shfl_sync(0xffffffffu, d, l);
        return;
      }
      z0 = __shfl_sync(0xffffffffu, zz, 31);
    }
#else
    (void)gas; (void)c; (void)t_first; (void)z; (void)dtau;
#endif
  }
};

// SOS_PROFILE: one warp per term, four terms per block; the 50-level gas tables of the block's terms sit in shared memory.
// All lanes of a warp run the same (uniform) bookkeeping and store the same values.  Arrays [nterm][PC_LEVELS].
#define PROF_WARPS 4
__global__ void __launch_bounds__(PROF_WARPS * 32)
k_profile(const ProfTermDev *__restrict__ terms, int nterm, const double *__restrict__ altabs, const double *__restrict__ tauabs,
          double *__restrict__ scratch, double *__restrict__ zprof, double *__restrict__ h, double *__restrict__ pcaer,
          double *__restrict__ pcmol, int *__restrict__ nt, int *__restrict__ ier)
{
  __shared__ double s_alt[PC_NLEV], s_tabs[PROF_WARPS][PC_NLEV];
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int i = blockIdx.x * PROF_WARPS + w;
  for (int k = threadIdx.x; k < PC_NLEV; k += blockDim.x) s_alt[k] = altabs[k];
  if (i < nterm) for (int k = lane; k < PC_NLEV; k += 32) s_tabs[w][k] = tauabs[(size_t)i * PC_NLEV + k];
  __syncthreads();
  if (i >= nterm) return;
  if (ier[i] != 0) { if (lane == 0) nt[i] = 0; return; }
  const ProfTermDev &t = terms[i];
  const size_t o = (size_t)i * PC_LEVELS;
  int n = 0;
  const int rc = pc_profile(PcWarp(), t.iprofil, t.tr, t.hr, t.ta, t.ha, t.zmin, t.zmax, t.absprofil, s_alt, s_tabs[w], scratch + o,
                            zprof + o, h + o, pcaer + o, pcmol + o, &n);
  __syncwarp();
  if (lane == 0) { ier[i] = rc; nt[i] = (rc == 0) ? n : 0; }
}

// The PROFIL_TMP text hop (format 20 written by SOS_PROFILE, read by SOS.F:511-516), elementwise over (term, level <= NT).
__global__ void k_profile_text(int nterm, const int *__restrict__ nt, double *__restrict__ zprof, double *__restrict__ h,
                               double *__restrict__ pcaer, double *__restrict__ pcmol)
{
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (size_t)nterm * PC_LEVELS) return;
  const int term = (int)(idx / PC_LEVELS), lev = (int)(idx - (size_t)term * PC_LEVELS);
  if (lev > nt[term]) { zprof[idx] = 0.0; h[idx] = 0.0; pcaer[idx] = 0.0; pcmol[idx] = 0.0; return; }
  zprof[idx] = pc_round_f5(zprof[idx]);
  h[idx] = pc_round_e8(h[idx]);
  pcaer[idx] = pc_round_e8(pcaer[idx]);
  pcmol[idx] = pc_round_e8(pcmol[idx]);
}

const size_t KI_SLICE = (size_t)PC_NTMAX * PC_NPMAX * PC_NAI * PC_NBABS;      // doubles of KDIS_KI per spectral interval
const size_t KH_SLICE = (size_t)PC_NTMAX * PC_NPMAX * PC_NCMAX * PC_NAI;      // ... of KDIS_KI_H2O

// Both stages for nterm terms.  ckd/atm null: tauabs_in [nterm][50] is uploaded instead of computed.
int run_chain(sosgpu_ctx *ctx, const sosgpu_ckd *ckd, const sosgpu_gas_profile *atm, const double *altabs_host,
              const sosgpu_profile_term *terms, int nterm, const double *tauabs_in, bool do_profile, int text_hop,
              double *tauabs_out, int *nt, double *zprof, double *h, double *pcaer, double *pcmol, int *ier)
{
  if (!ctx) return SOSGPU_ERR_NO_DEVICE;
  if (!terms || nterm < 1 || !ier) { ctx->err = "profile chain: bad arguments"; return SOSGPU_ERR_ARG; }
  CK(cudaSetDevice(ctx->device));
  cudaStream_t st = ctx->stream;
  SosFreeGuard guard(ctx);
  // ---- terms, with the spectral intervals they use packed densely ----
  std::vector<ProfTermDev> ht(nterm);
  std::map<int, int> slot;                                         // LAMB1 -> packed index (1-based)
  std::vector<int> lambs;
  for (int i = 0; i < nterm; ++i) {
    const sosgpu_profile_term &s = terms[i];
    ProfTermDev &d = ht[i];
    d.absprofil = s.absprofil; d.iprofil = s.iprofil;
    d.tr = s.tr; d.hr = s.hr; d.ta = s.ta; d.ha = s.ha; d.zmin = s.zmin; d.zmax = s.zmax;
    d.lamb = 1;
    for (int k = 0; k < PC_NBABS; ++k) d.ik[k] = s.ik[k];
    if (ckd && s.absprofil != 7) {
      if (s.lamb1 < 1 || s.lamb1 > PC_NWVL) { ctx->err = "profile chain: LAMB1 outside 1..50"; return SOSGPU_ERR_ARG; }
      for (int k = 0; k < PC_NBABS; ++k) {
        const int ne = ckd->nexp[PC_NEXP(k + 1, s.lamb1)];
        if (ne >= 1 && (s.ik[k] < 1 || s.ik[k] > PC_NAI)) { ctx->err = "profile chain: exponential index outside 1..5"; return SOSGPU_ERR_ARG; }
      }
      auto it = slot.find(s.lamb1);
      if (it == slot.end()) { lambs.push_back(s.lamb1); it = slot.emplace(s.lamb1, (int)lambs.size()).first; }
      d.lamb = it->second;
    }
  }
  ProfTermDev *d_terms = nullptr;
  int *d_ier = nullptr, *d_nt = nullptr;
  double *d_tau = nullptr;
  CK(sos_dmalloc(ctx, &d_terms, sizeof(ProfTermDev) * nterm)); guard.add(d_terms);
  CK(sos_dmalloc(ctx, &d_ier, sizeof(int) * nterm * 2)); guard.add(d_ier);
  d_nt = d_ier + nterm;
  CK(sos_dmalloc(ctx, &d_tau, sizeof(double) * nterm * PC_NLEV)); guard.add(d_tau);
  CK(cudaMemcpyAsync(d_terms, ht.data(), sizeof(ProfTermDev) * nterm, cudaMemcpyHostToDevice, st));
  CK(cudaMemsetAsync(d_ier, 0, sizeof(int) * nterm * 2, st));
  bool timed = false;                                              // ev_a .. ev_b bracket the kernels of this call
  std::vector<double> packed;                                      // lives until the stream has consumed it
  std::vector<int> nexp_packed;
  double *d_atm = nullptr;                                         // [altabs 50 | userprofil 650 | ro 400]
  CK(sos_dmalloc(ctx, &d_atm, sizeof(double) * (PC_NLEV + PC_NLEV * PC_NCOL + PC_NBABS * PC_NLEV))); guard.add(d_atm);
  if (altabs_host) {
    for (int k = 1; k < PC_NLEV; ++k)                              // the level search is a bisection of this table
      if (!(altabs_host[k] < altabs_host[k - 1])) { ctx->err = "profile chain: ALTABS must be strictly descending"; return SOSGPU_ERR_ARG; }
    CK(cudaMemcpyAsync(d_atm, altabs_host, sizeof(double) * PC_NLEV, cudaMemcpyHostToDevice, st));
  }
  if (ckd) {
    if (!atm || !atm->userprofil || !atm->ro || !ckd->tab_temp || !ckd->tab_pres || !ckd->tab_conc_h2o || !ckd->nexp || !ckd->kdis_ki ||
        !ckd->kdis_ki_h2o) { ctx->err = "profile chain: null table"; return SOSGPU_ERR_ARG; }
    if (ckd->nb_temp < 2 || ckd->nb_temp > PC_NTMAX || ckd->nb_pres < 2 || ckd->nb_pres > PC_NPMAX || ckd->nb_conc_h2o < 1 ||
        ckd->nb_conc_h2o > PC_NCMAX) { ctx->err = "profile chain: table extents outside inc/SOS.h's"; return SOSGPU_ERR_ARG; }
    CK(cudaMemcpyAsync(d_atm + PC_NLEV, atm->userprofil, sizeof(double) * PC_NLEV * PC_NCOL, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(d_atm + PC_NLEV + PC_NLEV * PC_NCOL, atm->ro, sizeof(double) * PC_NBABS * PC_NLEV, cudaMemcpyHostToDevice, st));
    const size_t nl = std::max<size_t>(lambs.size(), 1);
    const size_t head = PC_NTMAX + PC_NPMAX + PC_NCMAX;
    packed.assign(head + nl * (KI_SLICE + KH_SLICE), 0.0);
    memcpy(&packed[0], ckd->tab_temp, sizeof(double) * ckd->nb_temp);
    memcpy(&packed[PC_NTMAX], ckd->tab_pres, sizeof(double) * ckd->nb_pres);
    memcpy(&packed[PC_NTMAX + PC_NPMAX], ckd->tab_conc_h2o, sizeof(double) * ckd->nb_conc_h2o);
    nexp_packed.assign(nl * PC_NBABS, 0);
    for (size_t l = 0; l < lambs.size(); ++l) {
      const int lamb = lambs[l];
      memcpy(&packed[head + l * KI_SLICE], ckd->kdis_ki + (size_t)(lamb - 1) * KI_SLICE, sizeof(double) * KI_SLICE);
      memcpy(&packed[head + nl * KI_SLICE + l * KH_SLICE], ckd->kdis_ki_h2o + (size_t)(lamb - 1) * KH_SLICE, sizeof(double) * KH_SLICE);
      for (int k = 0; k < PC_NBABS; ++k) nexp_packed[l * PC_NBABS + k] = ckd->nexp[PC_NEXP(k + 1, lamb)];
    }
    double *d_packed = nullptr;
    int *d_nexp = nullptr;
    CK(sos_dmalloc(ctx, &d_packed, sizeof(double) * packed.size())); guard.add(d_packed);
    CK(sos_dmalloc(ctx, &d_nexp, sizeof(int) * nexp_packed.size())); guard.add(d_nexp);
    CK(cudaMemcpyAsync(d_packed, packed.data(), sizeof(double) * packed.size(), cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(d_nexp, nexp_packed.data(), sizeof(int) * nexp_packed.size(), cudaMemcpyHostToDevice, st));
    PcCkd c{};
    c.nb_temp = ckd->nb_temp; c.nb_pres = ckd->nb_pres; c.nb_conc = ckd->nb_conc_h2o;
    c.tab_temp = d_packed; c.tab_pres = d_packed + PC_NTMAX; c.tab_conc = d_packed + PC_NTMAX + PC_NPMAX;
    c.nexp = d_nexp; c.ki = d_packed + head; c.ki_h2o = d_packed + head + nl * KI_SLICE;
    if (ctx->ev_a) cudaEventRecord(ctx->ev_a, st);
    timed = true;
    k_absprofile<<<nterm, 64, 0, st>>>(c, d_atm + PC_NLEV, d_atm + PC_NLEV + PC_NLEV * PC_NCOL, d_terms, d_tau, d_ier);
    CK(cudaGetLastError());
    ctx->launches += 1;
  } else {
    if (!tauabs_in) { ctx->err = "profile chain: neither CKD tables nor an absorption profile"; return SOSGPU_ERR_ARG; }
    CK(cudaMemcpyAsync(d_tau, tauabs_in, sizeof(double) * nterm * PC_NLEV, cudaMemcpyHostToDevice, st));
  }
  if (tauabs_out) CK(cudaMemcpyAsync(tauabs_out, d_tau, sizeof(double) * nterm * PC_NLEV, cudaMemcpyDeviceToHost, st));
  double *d_prof = nullptr;
  const size_t per = (size_t)nterm * PC_LEVELS;
  if (do_profile) {
    if (!altabs_host || !nt || !zprof || !h || !pcaer || !pcmol) { ctx->err = "profile chain: null output"; return SOSGPU_ERR_ARG; }
    CK(sos_dmalloc(ctx, &d_prof, sizeof(double) * per * 5)); guard.add(d_prof);
    CK(cudaMemsetAsync(d_prof, 0, sizeof(double) * per * 5, st));
    if (!timed && ctx->ev_a) cudaEventRecord(ctx->ev_a, st);
    timed = true;
    k_profile<<<(nterm + PROF_WARPS - 1) / PROF_WARPS, PROF_WARPS * 32, 0, st>>>(d_terms, nterm, d_atm, d_tau, d_prof, d_prof + per, d_prof + 2 * per, d_prof + 3 * per,
                                                d_prof + 4 * per, d_nt, d_ier);
    CK(cudaGetLastError());
    ctx->launches += 1;
    if (text_hop) {
      k_profile_text<<<(unsigned)((per + 255) / 256), 256, 0, st>>>(nterm, d_nt, d_prof + per, d_prof + 2 * per, d_prof + 3 * per,
                                                                    d_prof + 4 * per);
      CK(cudaGetLastError());
      ctx->launches += 1;
    }
    if (ctx->ev_b) cudaEventRecord(ctx->ev_b, st);
    CK(cudaMemcpyAsync(zprof, d_prof + per, sizeof(double) * per, cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(h, d_prof + 2 * per, sizeof(double) * per, cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(pcaer, d_prof + 3 * per, sizeof(double) * per, cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(pcmol, d_prof + 4 * per, sizeof(double) * per, cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(nt, d_nt, sizeof(int) * nterm, cudaMemcpyDeviceToHost, st));
  }
  if (!do_profile && timed && ctx->ev_b) cudaEventRecord(ctx->ev_b, st);
  CK(cudaMemcpyAsync(ier, d_ier, sizeof(int) * nterm, cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st));
  if (timed && ctx->ev_a && ctx->ev_b) cudaEventElapsedTime(&ctx->last_kernel_ms, ctx->ev_a, ctx->ev_b);
  return SOSGPU_OK;
}

std::string fstr_(const char *s, size_t len)
{
  while (len > 0 && s[len - 1] == ' ') --len;
  return std::string(s, len);
}

sosgpu_ctx *profile_shim_ctx() { return sos_shim_ctx(); }

}  // namespace

extern "C" int sosgpu_absprofile(sosgpu_ctx *ctx, const sosgpu_ckd *ckd, const sosgpu_gas_profile *atm,
                                 const sosgpu_profile_term *terms, int nterm, double *tauabs, int *ier)
{
  if (!ctx) return SOSGPU_ERR_NO_DEVICE;
  if (!ckd || !atm || !tauabs) { ctx->err = "sosgpu_absprofile: bad arguments"; return SOSGPU_ERR_ARG; }
  return run_chain(ctx, ckd, atm, nullptr, terms, nterm, nullptr, false, 0, tauabs, nullptr, nullptr, nullptr, nullptr, nullptr, ier);
}

extern "C" int sosgpu_profile(sosgpu_ctx *ctx, const double *altabs, const double *tauabs, const sosgpu_profile_term *terms, int nterm,
                              int text_hop, int *nt, double *zprof, double *h, double *pcaer, double *pcmol, int *ier)
{
  if (!ctx) return SOSGPU_ERR_NO_DEVICE;
  if (!altabs || !tauabs) { ctx->err = "sosgpu_profile: bad arguments"; return SOSGPU_ERR_ARG; }
  return run_chain(ctx, nullptr, nullptr, altabs, terms, nterm, tauabs, true, text_hop, nullptr, nt, zprof, h, pcaer, pcmol, ier);
}

extern "C" int sosgpu_profile_chain(sosgpu_ctx *ctx, const sosgpu_ckd *ckd, const sosgpu_gas_profile *atm,
                                    const sosgpu_profile_term *terms, int nterm, int text_hop, double *tauabs, int *nt, double *zprof,
                                    double *h, double *pcaer, double *pcmol, int *ier)
{
  if (!ctx) return SOSGPU_ERR_NO_DEVICE;
  if (!ckd || !atm || !atm->altabs) { ctx->err = "sosgpu_profile_chain: bad arguments"; return SOSGPU_ERR_ARG; }
  return run_chain(ctx, ckd, atm, atm->altabs, terms, nterm, nullptr, true, text_hop, tauabs, nt, zprof, h, pcaer, pcmol, ier);
}

// ------------------------------------------------------------------------------------------------
// READ_CKD_COEFF (SOS_SUB_TRS.F:481-905): the CKD coefficients of gas NABS for the file that holds wavenumber NU.  Host code
// (it parses a text file): $SOS_ABS_ROOT/fic/COEFF_CKD/<step>cmm1/coef_<GAS>_<numax>_<numin>_<step>cmm1, 18 header lines (21
// for H2O), then list-directed numbers.  Arrays in the reference's Fortran storage.  Returns 0, or -1 as the reference's IER.
static bool next_number(FILE *f, double *v)
{
  char tok[128];
  int c = fgetc(f);
  while (c != EOF && (c == ' ' || c == '\t' || c == '\n' || c == '\r' || c == ',')) c = fgetc(f);
  if (c == EOF) return false;
  size_t n = 0;
  while (c != EOF && !(c == ' ' || c == '\t' || c == '\n' || c == '\r' || c == ',')) { if (n + 1 < sizeof tok) tok[n++] = (char)c; c = fgetc(f); }
  tok[n] = 0;
  for (size_t i = 0; i < n; ++i) if (tok[i] == 'D' || tok[i] == 'd') tok[i] = 'E';
  char *end = nullptr;
  *v = strtod(tok, &end);
  return end != tok;
}

extern "C" int sosgpu_read_ckd_coeff(const char *sos_abs_root, int nabs, int jabs, double nu, double nustep, int *nexp,
                                     double *kdis_ai, double *kdis_ki, double *kdis_ki_h2o, double *numax, double *numin,
                                     double *tab_pres, int *nb_pres, double *tab_temp, int *nb_temp, double *tab_conc_h2o,
                                     int *nb_conc_h2o)
{
  if (!nexp || !kdis_ai || !kdis_ki || !kdis_ki_h2o) return -1;
  if (nustep != 1 && nustep != 5 && nustep != 10) {
    printf("  SOS_SUB_TRS : ERROR_905: \n  The required spectral resolution is not supported :\n %g cm-1\n", nustep);
    return -1;
  }
  if (nabs < 1 || nabs > 8) { printf("  SOS_SUB_TRS : ERROR_900: \n  Index of gas not in the range [1,8]\n"); return -1; }
#define AI(nai, k, iwa) kdis_ai[((nai) - 1) + PC_NAI * (((k) - 1) + PC_NBABS * (size_t)((iwa) - 1))]
  if (jabs == 0) {                                                 // gas not selected: one exponential of weight 1, k = 0
    for (int iwa = 1; iwa <= PC_NWVL; ++iwa) {
      nexp[PC_NEXP(nabs, iwa)] = 1;
      AI(1, nabs, iwa) = 1.0;
      if (nabs == 1) {
        for (int ic = 1; ic <= PC_NCMAX; ++ic) for (int ip = 1; ip <= PC_NPMAX; ++ip) for (int it = 1; it <= PC_NTMAX; ++it)
          kdis_ki_h2o[PC_KI_H2O(it, ip, ic, 1, iwa)] = 0.0;
      } else {
        for (int ip = 1; ip <= PC_NPMAX; ++ip) for (int it = 1; it <= PC_NTMAX; ++it) kdis_ki[PC_KI(it, ip, 1, nabs, iwa)] = 0.0;
      }
    }
    return 0;
  }
  const char *root = sos_abs_root ? sos_abs_root : getenv("SOS_ABS_ROOT");
  if (!root || !*root) { printf("  SOS_SUB_TRS : ERROR_925: \n  => Error while getting SOS_ABS_ROOT variable\n"); return -1; }
  static const char *gas[8] = {"H2O", "CO2", "O3", "N2O", "CO", "CH4", "O2", "NO2"};
  const int step = (int)nustep, per_file = 50 * step;             // CTE_CKD_NB_NU_PER_FILE wavenumbers per file
  int numin_file = 27500 - per_file;                               // CTE_CKD_NUMAX
  while (numin_file > nu) numin_file -= per_file;
  const int numax_file = numin_file + per_file;
  char path[2048];
  snprintf(path, sizeof path, "%s/fic/COEFF_CKD/%dcmm1/coef_%s_%d_%d_%dcmm1", root, step, gas[nabs - 1], numax_file, numin_file, step);
  FILE *f = fopen(path, "r");
  if (!f) { printf("Error while opening the file of CKD coefficients\nFile :%s\n", path); return -1; }
  struct Closer { FILE *f; ~Closer() { fclose(f); } } closer{f};
  auto fail = [&](const char *what) { printf("Error while reading the file of CKD coefficients\nfile :%s\n%s\n", path, what); return -1; };
  char line[4096];
  for (int l = 0; l < (nabs == 1 ? 21 : 18); ++l) if (!fgets(line, sizeof line, f)) return fail("--> header");
  double v[8], resolution;
  if (!next_number(f, &v[0]) || !next_number(f, &v[1]) || !next_number(f, &resolution)) return fail("--> spectral range");
  if (resolution != nustep || v[0] != numax_file || v[1] != numin_file) {
    printf("Not consistent value of spectral range or resolution in\nfile :%s\n", path);
    return -1;
  }
  if (numax) *numax = v[0];
  if (numin) *numin = v[1];
  const int nb_wa = (int)((v[0] - v[1]) / resolution);
  if (nb_wa > PC_NWVL) { printf("Not consistent value of CTE_CKD_NWVL_MAX in fic/SOS.h\n"); return -1; }
  double x;
  if (!next_number(f, &x)) return fail("--> Block of temperature data");
  const int nt = (int)x;
  if (nt < 1 || nt > PC_NTMAX) return fail("--> Block of temperature data");
  for (int i = 0; i < nt; ++i) if (!next_number(f, &tab_temp[i])) return fail("--> Block of temperature data");
  if (!next_number(f, &x)) return fail("--> Block of pressure data");
  const int np = (int)x;
  if (np < 1 || np > PC_NPMAX) return fail("--> Block of pressure data");
  for (int i = 0; i < np; ++i) if (!next_number(f, &tab_pres[i])) return fail("--> Block of pressure data");
  *nb_temp = nt; *nb_pres = np;
  int nc = 1;
  if (nabs == 1) {
    if (!next_number(f, &x)) return fail("--> Block of pressure data");
    nc = (int)x;
    if (nc < 1 || nc > PC_NCMAX) return fail("--> Block of pressure data");
    for (int i = 0; i < nc; ++i) if (!next_number(f, &tab_conc_h2o[i])) return fail("--> Block of pressure data");
    *nb_conc_h2o = nc;
  }
  for (int iwa = 1; iwa <= nb_wa; ++iwa) {
    for (int i = 0; i < 6; ++i) if (!next_number(f, &v[i])) return fail("--> Spectral range information");
    const int nmax = (int)v[5];
    if (nmax < 0 || nmax > PC_NAI) return fail("--> Spectral range information");
    if (nmax == 0) {                                               // no absorption in this interval
      nexp[PC_NEXP(nabs, iwa)] = 1;
      AI(1, nabs, iwa) = 1.0;
      if (nabs == 1) {
        for (int ic = 1; ic <= nc; ++ic) for (int ip = 1; ip <= np; ++ip) for (int it = 1; it <= nt; ++it)
          kdis_ki_h2o[PC_KI_H2O(it, ip, ic, 1, iwa)] = 0.0;
      } else {
        for (int ip = 1; ip <= np; ++ip) for (int it = 1; it <= nt; ++it) kdis_ki[PC_KI(it, ip, 1, nabs, iwa)] = 0.0;
      }
      continue;
    }
    nexp[PC_NEXP(nabs, iwa)] = nmax;
    for (int nai = 1; nai <= nmax; ++nai) if (!next_number(f, &AI(nai, nabs, iwa))) return fail("--> Coeff ai");
    for (int nai = 1; nai <= nmax; ++nai) {
      for (int ic = 1; ic <= nc; ++ic) {
        for (int ip = 1; ip <= np; ++ip) {
          for (int i = 0; i < (nabs == 1 ? 3 : 2); ++i) if (!next_number(f, &x)) return fail("--> Coeff ki");   // NAI, [NIC,] NIP as read
          for (int it = 1; it <= nt; ++it) {
            double *dst = (nabs == 1) ? &kdis_ki_h2o[PC_KI_H2O(it, ip, ic, nai, iwa)] : &kdis_ki[PC_KI(it, ip, nai, nabs, iwa)];
            if (!next_number(f, dst)) return fail("--> Coeff ki");
          }
        }
      }
    }
  }
#undef AI
  return 0;
}

// ---- gfortran-ABI symbols (INTEGER*2 arguments are short) ----------------------------------------
// SOS_SUB_TRS.F:481-485
extern "C" void read_ckd_coeff_(const short *nabs, const short *jabs, const double *nu, const double *nustep, int *nexp, double *kdis_ai,
                                double *kdis_ki, double *kdis_ki_h2o, double *numax, double *numin, double *tab_pres, int *nb_pres,
                                double *tab_temp, int *nb_temp, double *tab_conc_h2o, int *nb_conc_h2o, int *ier)
{
  *ier = sosgpu_read_ckd_coeff(nullptr, *nabs, *jabs, *nu, *nustep, nexp, kdis_ai, kdis_ki, kdis_ki_h2o, numax, numin, tab_pres, nb_pres,
                               tab_temp, nb_temp, tab_conc_h2o, nb_conc_h2o) == 0 ? 0 : -1;
}

// SOS_ABSPROFILE.F:184-190.  The trace block (:385-401) is not written.
extern "C" void sos_absprofile_(const short *absprofil, const double *nu, const int *lamb1, const short *iabs, const double *userprofil,
                                const double *altabs, const double *ro, const int *nexp, const double *kdis_ki, const double *kdis_ki_h2o,
                                const int *ik1, const int *ik2, const int *ik3, const int *ik4, const int *ik5, const int *ik6,
                                const int *ik7, const int *ik8, const double *tab_pres, const int *nb_pres, const double *tab_temp,
                                const int *nb_temp, const double *tab_conc_h2o, const int *nb_conc_h2o, double *tauabstot,
                                const int *trace, const int *idlog, int *ier)
{
  (void)nu; (void)iabs; (void)trace; (void)idlog;
  *ier = 0;
  for (int j = 0; j < PC_NLEV; ++j) tauabstot[j] = 0.0;
  if (*absprofil == 7) return;
  sosgpu_ctx *ctx = profile_shim_ctx();
  if (!ctx) { printf("  SOS_ABSPROFILE : no usable CUDA device\n"); *ier = -1; return; }
  sosgpu_ckd c{};
  c.nb_temp = *nb_temp; c.nb_pres = *nb_pres; c.nb_conc_h2o = *nb_conc_h2o;
  c.tab_temp = tab_temp; c.tab_pres = tab_pres; c.tab_conc_h2o = tab_conc_h2o; c.nexp = nexp; c.kdis_ki = kdis_ki; c.kdis_ki_h2o = kdis_ki_h2o;
  sosgpu_gas_profile a{};
  a.userprofil = userprofil; a.altabs = altabs; a.ro = ro;
  sosgpu_profile_term t{};
  t.lamb1 = *lamb1; t.absprofil = *absprofil; t.iprofil = 1;
  const int *ik[8] = {ik1, ik2, ik3, ik4, ik5, ik6, ik7, ik8};
  for (int k = 0; k < 8; ++k) t.ik[k] = *ik[k];
  int e = 0;
  const int rc = sosgpu_absprofile(ctx, &c, &a, &t, 1, tauabstot, &e);
  if (rc != SOSGPU_OK || e != 0) {
    printf("  SOS_ABSPROFILE : ERROR_910 : \n  --> Error in subroutine COEFF_ABS_CKD\n");
    *ier = -1;
  }
}

// SOS_PROFIL.F:224-226 (hidden length of FICPROFIL last).  Writes FICPROFIL with format 20 (2X,I5,F10.5,3(E15.8)); the trace
// block is not written.
static void fmt_e15_8(char *out, double x)
{
  // Fortran E15.8: 0.dddddddd E+ee, right-justified in 15 columns
  char m[64], s[64];
  if (x == 0.0) snprintf(s, sizeof s, "0.00000000E+00");
  else {
    snprintf(m, sizeof m, "%.7E", fabs(x));                        // d.dddddddE+ee, correctly rounded
    char *e = strchr(m, 'E');
    const int ex = atoi(e + 1) + 1;
    *e = 0;
    char dig[16]; int nd = 0;
    for (char *c = m; *c; ++c) if (*c != '.') dig[nd++] = *c;
    dig[nd] = 0;
    if (ex <= -100 || ex >= 100) snprintf(s, sizeof s, "%s0.%s%c%03d", x < 0 ? "-" : "", dig, ex < 0 ? '-' : '+', abs(ex));
    else snprintf(s, sizeof s, "%s0.%sE%c%02d", x < 0 ? "-" : "", dig, ex < 0 ? '-' : '+', abs(ex));
  }
  if (strlen(s) > 15) snprintf(out, 16, "***************");
  else snprintf(out, 16, "%15s", s);
}

extern "C" void sos_profile_(const short *iprofil, const double *tr, const double *hr, const double *ta, const double *ha,
                             const double *zmin, const double *zmax, const short *absprofil, const double *altabs, const double *tabs,
                             const int *trace, const int *idlog, const char *ficprofil, int *nt, int *ier, size_t len_ficprofil)
{
  (void)trace; (void)idlog;
  *ier = 0;
  if (*iprofil != 1 && *iprofil != 2) {
    printf("  ERROR 940 in SOS_PROFIL :\n  IPROFIL must be 1 or 2\n  Current value :%d\n", (int)*iprofil);
    *ier = -1; return;
  }
  sosgpu_ctx *ctx = profile_shim_ctx();
  if (!ctx) { printf("  SOS_PROFIL : no usable CUDA device\n"); *ier = -1; return; }
  sosgpu_profile_term t{};
  t.lamb1 = 1; t.absprofil = *absprofil; t.iprofil = *iprofil;
  t.tr = *tr; t.hr = *hr; t.ta = *ta; t.ha = *ha; t.zmin = *zmin; t.zmax = *zmax;
  std::vector<double> z(PC_LEVELS), h(PC_LEVELS), pa(PC_LEVELS), pm(PC_LEVELS);
  int n = 0, e = 0;
  const int rc = sosgpu_profile(ctx, altabs, tabs, &t, 1, 0, &n, z.data(), h.data(), pa.data(), pm.data(), &e);
  if (rc != SOSGPU_OK || e != 0) {
    if (e == 1010) printf("  ERROR on boundary altitudes of the aerosol layer :\n  --> check -AP.AerLayer.Zmin and AP.AerLayer.Zmax\n");
    else if (e == 1020) printf("  AOT = %g is lower than the minimal value for the profile definition (0.00001)\n", *ta);
    else printf("  ERROR in SOS_PROFIL (%d)\n", rc != SOSGPU_OK ? rc : e);
    *ier = -1; return;
  }
  FILE *f = fopen(fstr_(ficprofil, len_ficprofil).c_str(), "w");
  if (!f) { printf("  ERROR on SOS_PROFIL result file opening\n"); *ier = -1; return; }
  for (int i = 0; i <= n; ++i) {
    char a[16], b[16], c[16], zf[32];
    fmt_e15_8(a, h[i]); fmt_e15_8(b, pa[i]); fmt_e15_8(c, pm[i]);
    snprintf(zf, sizeof zf, "%10.5f", z[i]);
    if (strlen(zf) > 10) snprintf(zf, sizeof zf, "**********");
    if (fprintf(f, "  %5d%s%s%s%s\n", i, zf, a, b, c) < 0) { fclose(f); printf("  ERROR on SOS_PROFIL result file writing\n"); *ier = -1; return; }
  }
  fclose(f);
  *nt = n;
}
