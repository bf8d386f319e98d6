// sosgpu_aerosols.cu -- aerosol optics per wavelength of libsosgpu.so: Mie tables (SOS_MIE, SOS_FPHASE_MIE),
// their integration over size distributions (SOS_GRANU), the mixture of modes (SOS_AEROSOLS) and the Legendre expansion with
// the optional truncation (SOS_DECOMPO_LEGENDRE) -- what the reference runs serially on the host, through MIE files, for every
// wavelength of a sweep and whose output are the alpha, beta, gamma, zeta coefficients the solver consumes.
//   k_mie_coef         one CTA per group of 32 size parameters (work list sorted by size parameter, all tables together): the lanes
//                      of a warp run the same serial recurrence for 32 size parameters, five recurrences on three warps, a_n / b_n
//                      and the sums over n strided over the warps; work arrays interleaved lane by lane (coalesced)
//   k_mie_phase        one warp per (group, scattering angle): the amplitude sums over n of 32 size parameters
//   k_granu            one CTA per component: weights of all records in parallel, sums over records one angle per thread
//   k_legendre_tables  P_k(mu_j), P^2_k(mu_j) once per angle set
//   k_model            one CTA per aerosol model: mixture, truncation, expansion one order per thread
// The arithmetic is csrc/aerosol_chain.cuh (statement order of the reference; compiled with -fmad=false), so results differ from
// the reference's only through the device's sin / cos / exp / log / pow / acos (<= 2 ulp) -- REAL*4 Mie records are bit-identical
// except where such a difference crosses a single-precision rounding boundary.
// Also here: the gfortran-ABI symbols sos_mie_, sos_granu_, sos_decompo_legendre_ and the writer of the aerosol result file.
#include "sosgpu_host.h"
#include "aerosol_chain.cuh"
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <algorithm>
#include <map>
#include <tuple>

namespace {

struct AcCtaSync { __device__ void operator()() const { __syncthreads(); } };

struct MieTableDev { double rn, in; long long rec0; int nrec, pad; };
struct MieItem { int table, rec; };
struct MieGroupDev { long long off; int stride, pad; };
struct GranuCompDev { long long rec0; int nrec, igranu; double alphaf, v1, v2, v3, wa; };

constexpr int COEF_WARPS = 4;

// stage 1: blockIdx.x = group of the chunk that starts at group0
__global__ void __launch_bounds__(COEF_WARPS * 32, 4)
k_mie_coef(const MieTableDev *__restrict__ tables, const MieItem *__restrict__ items, const MieGroupDev *__restrict__ groups, int group0,
           const double *__restrict__ alpha, double *__restrict__ arena, float *__restrict__ rec, double *__restrict__ g,
           int *__restrict__ it_n2, double *__restrict__ it_qsca)
{
  __shared__ AcLaneState st[32];
  const int grp = group0 + (int)blockIdx.x, lane = (int)threadIdx.x & 31, wr = (int)threadIdx.x >> 5;
  const size_t it = (size_t)grp * 32 + lane;
  const MieItem wi = items[it];
  const bool valid = wi.table >= 0;
  const MieGroupDev gd = groups[grp];
  const AcMieWork w = ac_work(arena + gd.off + lane, (size_t)gd.stride, 32);
  MieTableDev t{};
  size_t r = 0;
  double a = 1.0;
  if (valid) { t = tables[wi.table]; r = (size_t)t.rec0 + wi.rec; a = alpha[r]; }
  for (int ph = 0; ph < AC_COEF_PHASES; ++ph) {
    if (valid) ac_coef_phase(ph, wr, COEF_WARPS, a, t.rn, t.in, w, st[lane], rec + 3 * r, g + r, it_n2 + it, it_qsca + it);
    __syncthreads();
  }
}

// stage 2: one warp per (group, angle) of the chunk; tasks of the heaviest groups first
__global__ void __launch_bounds__(128)
k_mie_phase(const MieTableDev *__restrict__ tables, const MieItem *__restrict__ items, const MieGroupDev *__restrict__ groups, int group0,
            int ngroups, int nang, const double *__restrict__ rmu, const double *__restrict__ alpha, const double *__restrict__ arena,
            const int *__restrict__ it_n2, const double *__restrict__ it_qsca, float *__restrict__ imie, float *__restrict__ qmie,
            float *__restrict__ umie)
{
  const long long task = (long long)blockIdx.x * 4 + (threadIdx.x >> 5);
  if (task >= (long long)ngroups * nang) return;
  const int grp = group0 + (int)(task / nang), j = (int)(task % nang), lane = (int)threadIdx.x & 31;
  const size_t it = (size_t)grp * 32 + lane;
  const MieItem wi = items[it];
  if (wi.table < 0) return;
  const MieGroupDev gd = groups[grp];
  const AcMieWork w = ac_work(const_cast<double *>(arena) + gd.off + lane, (size_t)gd.stride, 32);
  const size_t r = (size_t)tables[wi.table].rec0 + wi.rec;
  const AcCoef cf{w.ra, w.ia, w.rb, w.ib, w.es};
  const size_t o = r * (size_t)nang + j;
  ac_mie_phase(rmu[j], alpha[r], it_qsca[it], it_n2[it], cf, imie + o, qmie + o, umie + o);
}

__global__ void __launch_bounds__(128)
k_granu(const GranuCompDev *__restrict__ comps, int nang, const float *__restrict__ rec, const float *__restrict__ imie,
        const float *__restrict__ qmie, const float *__restrict__ umie, double *__restrict__ scratch, size_t scratch_stride,
        double *__restrict__ comp_k, double *__restrict__ p11, double *__restrict__ p12, double *__restrict__ p33, int *__restrict__ ier)
{
  __shared__ int sh_k;
  const GranuCompDev c = comps[blockIdx.x];
  const size_t r0 = (size_t)c.rec0, o = (size_t)blockIdx.x * nang;
  ac_granu((int)threadIdx.x, (int)blockDim.x, AcCtaSync(), c.nrec, rec + 3 * r0, imie + r0 * nang, qmie + r0 * nang, umie + r0 * nang, nang,
           c.alphaf, c.igranu, c.v1, c.v2, c.v3, c.wa, scratch + (size_t)blockIdx.x * scratch_stride, &sh_k, comp_k + 3 * (size_t)blockIdx.x,
           p11 + o, p12 + o, p33 + o, ier + blockIdx.x);
}

__global__ void k_legendre_tables(int nang, const double *__restrict__ xmu, int nb, double *__restrict__ pl, double *__restrict__ pol)
{
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j < nang) ac_legendre_column(xmu[j], nb, pl + j, pol + j, (size_t)nang);
}

__global__ void __launch_bounds__(256)
k_model(int nbmu, const double *__restrict__ xmu, const double *__restrict__ xhr, const double *__restrict__ pl, const double *__restrict__ pol,
        const double *__restrict__ comp_k, const double *__restrict__ p11, const double *__restrict__ p12, const double *__restrict__ p33,
        const double *__restrict__ p22, const AcModel *__restrict__ models, double *__restrict__ scal, double *__restrict__ coef,
        double *__restrict__ phase, int *__restrict__ ier)
{
  __shared__ AcModelShared s;
  const AcModel m = models[blockIdx.x];
  const size_t nang = 2 * (size_t)nbmu + 1;
  ac_model((int)threadIdx.x, (int)blockDim.x, AcCtaSync(), nbmu, xmu, xhr, pl, pol, comp_k, p11, p12, p33, p22, m, s,
           scal + 8 * (size_t)blockIdx.x, coef + (size_t)blockIdx.x * 6 * (m.os_nb + 1), phase ? phase + (size_t)blockIdx.x * 4 * nang : nullptr,
           ier + blockIdx.x);
}

// ---- host side ----------------------------------------------------------------------------------
// the size-parameter grid of SOS_MIE (SOS_MIE.F:399-411, 651-652); false: its error 997
bool mie_grid(double alpha0, double alphaf, std::vector<double> *out)
{
  if (trunc(alphaf + alphaf + 20) > AC_MIE_DIM) return false;
  if (!(alpha0 > 0.0)) return false;
  for (double a = alpha0;;) {
    if (out) out->push_back(a);
    a = a + ac_mie_step(a);
    if (!(a <= alphaf)) break;
  }
  return true;
}

struct Arena {                 // one pool allocation carved into aligned pieces
  size_t bytes = 0;
  size_t take(size_t n) { const size_t o = bytes; bytes += (n + 255) & ~(size_t)255; return o; }
};

struct HostTable { double rn, in, alpha0, alphaf; size_t rec0; int nrec; };

// Mie tables already on the device (from k_mie or uploaded from a file)
struct DevTables {
  float *rec = nullptr, *imie = nullptr, *qmie = nullptr, *umie = nullptr;
  double *g = nullptr;
};

// ev_a is recorded right before the first kernel: the stream is idle while the host allocates, so an earlier record would time
// the growth of the memory pool as well
constexpr size_t MIE_ARENA_DOUBLES = (size_t)1 << 28;      // 2 GB of work arrays per chunk of the work list

int launch_mie(sosgpu_ctx *ctx, int nbmu, const double *d_rmu, const std::vector<HostTable> &tabs, const std::vector<double> &alpha,
               const DevTables &dt, SosFreeGuard &guard)
{
  cudaStream_t st = ctx->stream;
  const size_t total = alpha.size();
  const int nang = 2 * nbmu + 1;
  std::vector<MieTableDev> td(tabs.size());
  std::vector<long long> rec0(tabs.size());
  std::vector<int> nrec(tabs.size());
  for (size_t t = 0; t < tabs.size(); ++t) {
    td[t] = MieTableDev{tabs[t].rn, tabs[t].in, (long long)tabs[t].rec0, tabs[t].nrec, 0};
    rec0[t] = (long long)tabs[t].rec0; nrec[t] = tabs[t].nrec;
  }
  const AcMiePlan plan = ac_mie_plan(rec0, nrec, alpha, MIE_ARENA_DOUBLES);
  const size_t nitem = plan.item_table.size(), ngroup = plan.group_off.size();
  std::vector<MieItem> items(nitem);
  for (size_t i = 0; i < nitem; ++i) items[i] = MieItem{plan.item_table[i], plan.item_rec[i]};
  std::vector<MieGroupDev> gd(ngroup);
  for (size_t g = 0; g < ngroup; ++g) gd[g] = MieGroupDev{plan.group_off[g], plan.group_stride[g], 0};
  Arena a;
  const size_t o_tab = a.take(sizeof(MieTableDev) * td.size()), o_it = a.take(sizeof(MieItem) * nitem), o_gd = a.take(sizeof(MieGroupDev) * ngroup);
  const size_t o_al = a.take(sizeof(double) * total), o_n2 = a.take(sizeof(int) * nitem), o_q = a.take(sizeof(double) * nitem);
  const size_t o_work = a.take(sizeof(double) * plan.arena);
  char *d = nullptr;
  CK(sos_dmalloc(ctx, &d, a.bytes)); guard.add(d);
  CK(cudaMemcpyAsync(d + o_tab, td.data(), sizeof(MieTableDev) * td.size(), cudaMemcpyHostToDevice, st));
  CK(cudaMemcpyAsync(d + o_it, items.data(), sizeof(MieItem) * nitem, cudaMemcpyHostToDevice, st));
  CK(cudaMemcpyAsync(d + o_gd, gd.data(), sizeof(MieGroupDev) * ngroup, cudaMemcpyHostToDevice, st));
  CK(cudaMemcpyAsync(d + o_al, alpha.data(), sizeof(double) * total, cudaMemcpyHostToDevice, st));
  if (ctx->ev_a) cudaEventRecord(ctx->ev_a, st);
  for (size_t c = 0; c + 1 < plan.chunk_first.size(); ++c) {          // the arena is reused chunk after chunk, in stream order
    const int g0 = plan.chunk_first[c], ng = plan.chunk_first[c + 1] - g0;
    if (ng <= 0) continue;
    k_mie_coef<<<ng, COEF_WARPS * 32, 0, st>>>((const MieTableDev *)(d + o_tab), (const MieItem *)(d + o_it), (const MieGroupDev *)(d + o_gd), g0,
                                               (const double *)(d + o_al), (double *)(d + o_work), dt.rec, dt.g, (int *)(d + o_n2),
                                               (double *)(d + o_q));
    CK(cudaGetLastError());
    const long long tasks = (long long)ng * nang;
    k_mie_phase<<<(unsigned)((tasks + 3) / 4), 128, 0, st>>>((const MieTableDev *)(d + o_tab), (const MieItem *)(d + o_it),
                                                             (const MieGroupDev *)(d + o_gd), g0, ng, nang, d_rmu, (const double *)(d + o_al),
                                                             (const double *)(d + o_work), (const int *)(d + o_n2),
                                                             (const double *)(d + o_q), dt.imie, dt.qmie, dt.umie);
    CK(cudaGetLastError());
    ctx->launches += 2;
  }
  CK(cudaStreamSynchronize(st));             // the host vectors above are consumed
  return SOSGPU_OK;
}

int alloc_tables(sosgpu_ctx *ctx, size_t total, size_t nang, DevTables *dt, SosFreeGuard &guard)
{
  Arena a;
  const size_t o_rec = a.take(sizeof(float) * 3 * total), o_g = a.take(sizeof(double) * total);
  const size_t o_i = a.take(sizeof(float) * total * nang), o_q = a.take(sizeof(float) * total * nang), o_u = a.take(sizeof(float) * total * nang);
  char *d = nullptr;
  CK(sos_dmalloc(ctx, &d, a.bytes)); guard.add(d);
  dt->rec = (float *)(d + o_rec); dt->g = (double *)(d + o_g);
  dt->imie = (float *)(d + o_i); dt->qmie = (float *)(d + o_q); dt->umie = (float *)(d + o_u);
  return SOSGPU_OK;
}

bool angles_ok(sosgpu_ctx *ctx, int nbmu, const double *a, const char *who)
{
  if (nbmu < 1 || nbmu > AC_MIE_NBMU_MAX || !a) { ctx->err = std::string(who) + ": bad angle arguments (1 <= nbmu <= 100)"; return false; }
  return true;
}

std::string fstr(const char *s, size_t len)
{
  while (len > 0 && s[len - 1] == ' ') --len;
  return std::string(s, len);
}

sosgpu_ctx *shim_ctx() { return sos_shim_ctx(); }

}  // namespace

extern "C" int sosgpu_mie_count(double alpha0, double alphaf)
{
  std::vector<double> a;
  if (!mie_grid(alpha0, alphaf, &a)) return -1;
  return (int)a.size();
}

// SOS_MIE (SOS_MIE.F:205-690) without its file: the records of one table in host arrays
extern "C" int sosgpu_mie(sosgpu_ctx *ctx, int nbmu, const double *rmu, double rn, double in, double alpha0, double alphaf, int nrec_cap,
                          float *rec, double *g, float *imie, float *qmie, float *umie, int *nrec)
{
  if (!ctx) return SOSGPU_ERR_NO_DEVICE;
  if (!angles_ok(ctx, nbmu, rmu, "sosgpu_mie")) return SOSGPU_ERR_ARG;
  if (!rec || !g || !imie || !qmie || !umie || !nrec) { ctx->err = "sosgpu_mie: null output"; return SOSGPU_ERR_ARG; }
  std::vector<double> alpha;
  if (alpha0 > alphaf || !mie_grid(alpha0, alphaf, &alpha)) { ctx->err = "sosgpu_mie: size-parameter range outside CTE_MIE_DIM (SOS_MIE error 997)"; return SOSGPU_ERR_IER; }
  *nrec = (int)alpha.size();
  if ((int)alpha.size() > nrec_cap) { ctx->err = "sosgpu_mie: output capacity too small"; return SOSGPU_ERR_ARG; }
  CK(cudaSetDevice(ctx->device));
  cudaStream_t st = ctx->stream;
  const size_t nang = 2 * (size_t)nbmu + 1, total = alpha.size();
  SosFreeGuard guard(ctx);
  DevTables dt;
  int rc = alloc_tables(ctx, total, nang, &dt, guard);
  if (rc != SOSGPU_OK) return rc;
  double *d_rmu = nullptr;
  CK(sos_dmalloc(ctx, &d_rmu, sizeof(double) * nang)); guard.add(d_rmu);
  CK(cudaMemcpyAsync(d_rmu, rmu, sizeof(double) * nang, cudaMemcpyHostToDevice, st));
  std::vector<HostTable> tabs{HostTable{rn, in, alpha0, alphaf, 0, (int)total}};
  rc = launch_mie(ctx, nbmu, d_rmu, tabs, alpha, dt, guard);
  if (rc != SOSGPU_OK) return rc;
  if (ctx->ev_b) cudaEventRecord(ctx->ev_b, st);
  CK(cudaMemcpyAsync(rec, dt.rec, sizeof(float) * 3 * total, cudaMemcpyDeviceToHost, st));
  CK(cudaMemcpyAsync(g, dt.g, sizeof(double) * total, cudaMemcpyDeviceToHost, st));
  CK(cudaMemcpyAsync(imie, dt.imie, sizeof(float) * total * nang, cudaMemcpyDeviceToHost, st));
  CK(cudaMemcpyAsync(qmie, dt.qmie, sizeof(float) * total * nang, cudaMemcpyDeviceToHost, st));
  CK(cudaMemcpyAsync(umie, dt.umie, sizeof(float) * total * nang, cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st));
  if (ctx->ev_a && ctx->ev_b) cudaEventElapsedTime(&ctx->last_kernel_ms, ctx->ev_a, ctx->ev_b);
  return SOSGPU_OK;
}

namespace {

// components + models on tables that are on the device; host outputs may be null
int run_granu_models(sosgpu_ctx *ctx, int nbmu, const double *xmu, const double *xhr, const DevTables &dt, const std::vector<GranuCompDev> &comps,
                     const double *host_comp_in /* [ncomp][3 + 4 nang]: k, p11, p12, p33, p22 given instead of k_granu, or null */,
                     bool with_p22, const std::vector<AcModel> &models, int os_nb, double *comp_k, double *comp_phase, int *comp_ier, double *scal,
                     double *coef, double *phase, int *model_ier, SosFreeGuard &guard, bool start_timer = true)
{
  cudaStream_t st = ctx->stream;
  const size_t nang = 2 * (size_t)nbmu + 1, nc = comps.size(), nm = models.size();
  size_t maxrec = 1;
  for (const GranuCompDev &c : comps) maxrec = std::max<size_t>(maxrec, (size_t)c.nrec);
  const size_t sstride = 3 * maxrec + 8;
  Arena a;
  const size_t o_xmu = a.take(sizeof(double) * 2 * nang);
  const size_t o_comp = a.take(sizeof(GranuCompDev) * std::max<size_t>(nc, 1));
  const size_t o_scr = a.take(host_comp_in ? 8 : sizeof(double) * sstride * nc);
  const size_t o_ck = a.take(sizeof(double) * 3 * nc), o_p = a.take(sizeof(double) * 4 * nc * nang), o_cier = a.take(sizeof(int) * nc);
  const size_t o_pl = a.take(sizeof(double) * 2 * (os_nb + 1) * nang);
  const size_t o_mod = a.take(sizeof(AcModel) * std::max<size_t>(nm, 1)), o_scal = a.take(sizeof(double) * 8 * nm);
  const size_t o_coef = a.take(sizeof(double) * 6 * (os_nb + 1) * nm), o_ph = a.take(sizeof(double) * 4 * nang * nm), o_mier = a.take(sizeof(int) * nm);
  char *d = nullptr;
  CK(sos_dmalloc(ctx, &d, a.bytes)); guard.add(d);
  double *d_xmu = (double *)(d + o_xmu), *d_xhr = d_xmu + nang;
  double *d_ck = (double *)(d + o_ck), *d_p11 = (double *)(d + o_p), *d_p12 = d_p11 + nc * nang, *d_p33 = d_p12 + nc * nang, *d_p22 = d_p33 + nc * nang;
  double *d_pl = (double *)(d + o_pl), *d_pol = d_pl + (size_t)(os_nb + 1) * nang;
  int *d_cier = (int *)(d + o_cier), *d_mier = (int *)(d + o_mier);
  CK(cudaMemcpyAsync(d_xmu, xmu, sizeof(double) * nang, cudaMemcpyHostToDevice, st));
  if (xhr) CK(cudaMemcpyAsync(d_xhr, xhr, sizeof(double) * nang, cudaMemcpyHostToDevice, st));
  CK(cudaMemsetAsync(d_cier, 0xff, sizeof(int) * nc, st));           // -1 until a kernel says otherwise
  if (nm) CK(cudaMemsetAsync(d_mier, 0xff, sizeof(int) * nm, st));
  std::vector<double> stage;
  if (host_comp_in) {
    stage.resize(3 * nc + 4 * nc * nang);
    for (size_t c = 0; c < nc; ++c) {
      const double *src = host_comp_in + c * (3 + 4 * nang);
      memcpy(&stage[3 * c], src, sizeof(double) * 3);
      for (int q = 0; q < 4; ++q) memcpy(&stage[3 * nc + (q * nc + c) * nang], src + 3 + q * nang, sizeof(double) * nang);
    }
    CK(cudaMemcpyAsync(d_ck, stage.data(), sizeof(double) * 3 * nc, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(d_p11, stage.data() + 3 * nc, sizeof(double) * 4 * nc * nang, cudaMemcpyHostToDevice, st));
    CK(cudaMemsetAsync(d_cier, 0, sizeof(int) * nc, st));
    if (start_timer && ctx->ev_a) cudaEventRecord(ctx->ev_a, st);
  } else {
    CK(cudaMemcpyAsync(d + o_comp, comps.data(), sizeof(GranuCompDev) * nc, cudaMemcpyHostToDevice, st));
    if (start_timer && ctx->ev_a) cudaEventRecord(ctx->ev_a, st);
    k_granu<<<(unsigned)nc, 128, 0, st>>>((const GranuCompDev *)(d + o_comp), (int)nang, dt.rec, dt.imie, dt.qmie, dt.umie, (double *)(d + o_scr),
                                          sstride, d_ck, d_p11, d_p12, d_p33, d_cier);
    CK(cudaGetLastError());
    ctx->launches += 1;
  }
  if (nm) {
    k_legendre_tables<<<(unsigned)((nang + 63) / 64), 64, 0, st>>>((int)nang, d_xmu, os_nb, d_pl, d_pol);
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(d + o_mod, models.data(), sizeof(AcModel) * nm, cudaMemcpyHostToDevice, st));
    k_model<<<(unsigned)nm, 256, 0, st>>>(nbmu, d_xmu, d_xhr, d_pl, d_pol, d_ck, d_p11, d_p12, d_p33, with_p22 ? d_p22 : nullptr,
                                          (const AcModel *)(d + o_mod), (double *)(d + o_scal), (double *)(d + o_coef),
                                          (double *)(d + o_ph), d_mier);
    CK(cudaGetLastError());
    ctx->launches += 2;
  }
  if (ctx->ev_b) cudaEventRecord(ctx->ev_b, st);
  if (comp_k) CK(cudaMemcpyAsync(comp_k, d_ck, sizeof(double) * 3 * nc, cudaMemcpyDeviceToHost, st));
  std::vector<double> back;
  if (comp_phase) { back.resize(3 * nc * nang); CK(cudaMemcpyAsync(back.data(), d_p11, sizeof(double) * 3 * nc * nang, cudaMemcpyDeviceToHost, st)); }
  if (comp_ier) CK(cudaMemcpyAsync(comp_ier, d_cier, sizeof(int) * nc, cudaMemcpyDeviceToHost, st));
  if (nm) {
    if (scal) CK(cudaMemcpyAsync(scal, d + o_scal, sizeof(double) * 8 * nm, cudaMemcpyDeviceToHost, st));
    if (coef) CK(cudaMemcpyAsync(coef, d + o_coef, sizeof(double) * 6 * (os_nb + 1) * nm, cudaMemcpyDeviceToHost, st));
    if (phase) CK(cudaMemcpyAsync(phase, d + o_ph, sizeof(double) * 4 * nang * nm, cudaMemcpyDeviceToHost, st));
    if (model_ier) CK(cudaMemcpyAsync(model_ier, d_mier, sizeof(int) * nm, cudaMemcpyDeviceToHost, st));
  }
  CK(cudaStreamSynchronize(st));
  if (comp_phase)                                                   // device [quantity][component][angle] -> [component][quantity][angle]
    for (size_t c = 0; c < nc; ++c)
      for (int q = 0; q < 3; ++q) memcpy(comp_phase + (c * 3 + q) * nang, &back[(q * nc + c) * nang], sizeof(double) * nang);
  if (ctx->ev_a && ctx->ev_b) cudaEventElapsedTime(&ctx->last_kernel_ms, ctx->ev_a, ctx->ev_b);
  return SOSGPU_OK;
}

bool models_ok(sosgpu_ctx *ctx, const sosgpu_aer_model *models, int nmodel, int ncomp, int os_nb, std::vector<AcModel> *out)
{
  if (nmodel < 0 || (nmodel && !models) || os_nb < 2 || os_nb > AC_NB_MAX) { ctx->err = "aerosol models: bad arguments (2 <= os_nb <= 200)"; return false; }
  out->resize(nmodel);
  for (int m = 0; m < nmodel; ++m) {
    AcModel &d = (*out)[m];
    d = AcModel{};
    d.ncomp = models[m].ncomp; d.itronc = models[m].itronc; d.os_nb = os_nb;
    if (d.ncomp < 0 || d.ncomp > 4 || (d.itronc != 0 && d.itronc != 1)) { ctx->err = "aerosol models: ncomp outside 0..4 or itronc not 0 / 1"; return false; }
    for (int i = 0; i < std::max(d.ncomp, 1); ++i) {
      d.comp[i] = models[m].comp[i]; d.w[i] = models[m].weight[i];
      if (d.comp[i] < 0 || d.comp[i] >= ncomp) { ctx->err = "aerosol models: component index out of range"; return false; }
    }
  }
  return true;
}

}  // namespace

// SOS_GRANU (SOS_AEROSOLS.F:4392-4767) on a Mie table given in host arrays (what the reference reads back from its MIE file)
extern "C" int sosgpu_granu(sosgpu_ctx *ctx, int nbmu, int nrec, const float *rec, const float *imie, const float *qmie, const float *umie,
                            double alphaf, int igranu, double v1, double v2, double v3, double wa, double *kmat, double *p11, double *p12,
                            double *p33, int *ier)
{
  if (!ctx) return SOSGPU_ERR_NO_DEVICE;
  if (nbmu < 1 || nbmu > AC_MIE_NBMU_MAX || nrec < 1 || !rec || !imie || !qmie || !umie || !kmat || !p11 || !p12 || !p33 || !ier ||
      (igranu != 1 && igranu != 2)) { ctx->err = "sosgpu_granu: bad arguments"; return SOSGPU_ERR_ARG; }
  CK(cudaSetDevice(ctx->device));
  cudaStream_t st = ctx->stream;
  const size_t nang = 2 * (size_t)nbmu + 1, total = (size_t)nrec;
  SosFreeGuard guard(ctx);
  DevTables dt;
  int rc = alloc_tables(ctx, total, nang, &dt, guard);
  if (rc != SOSGPU_OK) return rc;
  CK(cudaMemcpyAsync(dt.rec, rec, sizeof(float) * 3 * total, cudaMemcpyHostToDevice, st));
  CK(cudaMemcpyAsync(dt.imie, imie, sizeof(float) * total * nang, cudaMemcpyHostToDevice, st));
  CK(cudaMemcpyAsync(dt.qmie, qmie, sizeof(float) * total * nang, cudaMemcpyHostToDevice, st));
  CK(cudaMemcpyAsync(dt.umie, umie, sizeof(float) * total * nang, cudaMemcpyHostToDevice, st));
  std::vector<GranuCompDev> comps{GranuCompDev{0, nrec, igranu, alphaf, v1, v2, v3, wa}};
  std::vector<double> zero(nang, 0.0), ph(3 * nang);
  rc = run_granu_models(ctx, nbmu, zero.data(), nullptr, dt, comps, nullptr, false, {}, 2, kmat, ph.data(), ier, nullptr, nullptr, nullptr, nullptr, guard);
  if (rc != SOSGPU_OK) return rc;
  memcpy(p11, &ph[0], sizeof(double) * nang); memcpy(p12, &ph[nang], sizeof(double) * nang); memcpy(p33, &ph[2 * nang], sizeof(double) * nang);
  return SOSGPU_OK;
}

// SOS_DECOMPO_LEGENDRE (SOS_AEROSOLS.F:3924-4260).  p11 in/out (truncated on exit), ttt out; the six coefficient arrays (0:os_nb)
// are overwritten (the reference accumulates into arrays its caller has zeroed, SOS_AEROSOLS.F:1113-1120).
extern "C" int sosgpu_decompo_legendre(sosgpu_ctx *ctx, int *itronc, int nbmu, const double *xmu, const double *xhr, int os_nb, double *p11,
                                       double *ttt, const double *p12, const double *p22, const double *p33, double *coef_tronca, double *z1,
                                       double *alp, double *beta11, double *beta22, double *gamma12, double *delta33, double *zeta, int *ier)
{
  if (!ctx) return SOSGPU_ERR_NO_DEVICE;
  if (!angles_ok(ctx, nbmu, xmu, "sosgpu_decompo_legendre")) return SOSGPU_ERR_ARG;
  if (!xhr || !itronc || !p11 || !p12 || !p22 || !p33 || !ier) { ctx->err = "sosgpu_decompo_legendre: null argument"; return SOSGPU_ERR_ARG; }
  sosgpu_aer_model mod{};
  mod.ncomp = 0; mod.comp[0] = 0; mod.weight[0] = 1.0; mod.itronc = *itronc;
  std::vector<AcModel> models;
  if (!models_ok(ctx, &mod, 1, 1, os_nb, &models)) return SOSGPU_ERR_ARG;
  CK(cudaSetDevice(ctx->device));
  const size_t nang = 2 * (size_t)nbmu + 1;
  std::vector<double> in(3 + 4 * nang, 1.0);                        // KMAT1 = KMAT2 = 1: not used by the expansion
  memcpy(&in[3], p11, sizeof(double) * nang); memcpy(&in[3 + nang], p12, sizeof(double) * nang);
  memcpy(&in[3 + 2 * nang], p33, sizeof(double) * nang); memcpy(&in[3 + 3 * nang], p22, sizeof(double) * nang);
  std::vector<GranuCompDev> comps(1);
  std::vector<double> scal(8), coef(6 * (size_t)(os_nb + 1)), ph(4 * nang);
  SosFreeGuard guard(ctx);
  const int rc = run_granu_models(ctx, nbmu, xmu, xhr, DevTables{}, comps, in.data(), true, models, os_nb, nullptr, nullptr, nullptr, scal.data(),
                                  coef.data(), ph.data(), ier, guard);
  if (rc != SOSGPU_OK) return rc;
  if (*ier != 0) return SOSGPU_OK;
  const size_t n = (size_t)os_nb + 1;
  memcpy(p11, &ph[0], sizeof(double) * nang);
  if (ttt) memcpy(ttt, &ph[3 * nang], sizeof(double) * nang);
  if (coef_tronca) *coef_tronca = scal[4];
  if (z1) *z1 = scal[6];
  *itronc = (int)scal[7];
  double *dst[6] = {alp, beta11, gamma12, zeta, beta22, delta33};
  for (int q = 0; q < 6; ++q) if (dst[q]) memcpy(dst[q], &coef[q * n], sizeof(double) * n);
  return SOSGPU_OK;
}

// The whole chain for a list of components (one per mode and wavelength) and models (one per wavelength): Mie tables -- one per
// distinct (rn, in, alpha0, alphaf), as the reference's MIE file cache -- size-distribution integrals, mixtures, expansions; the
// tables never leave the device.
extern "C" int sosgpu_aerosols(sosgpu_ctx *ctx, int nbmu, const double *xmu, const double *xhr, int ncomp, const sosgpu_aer_component *comp,
                               int nmodel, const sosgpu_aer_model *models, int os_nb, double *comp_k, double *comp_phase, int *comp_ier,
                               double *scal, double *coef, double *phase, int *model_ier)
{
  if (!ctx) return SOSGPU_ERR_NO_DEVICE;
  if (!angles_ok(ctx, nbmu, xmu, "sosgpu_aerosols")) return SOSGPU_ERR_ARG;
  if (!xhr || ncomp < 1 || !comp) { ctx->err = "sosgpu_aerosols: null weights or no component"; return SOSGPU_ERR_ARG; }
  std::vector<AcModel> mods;
  if (!models_ok(ctx, models, nmodel, ncomp, os_nb, &mods)) return SOSGPU_ERR_ARG;
  std::vector<HostTable> tabs;
  std::vector<double> alpha;
  std::map<std::tuple<double, double, double, double>, int> seen;
  std::vector<GranuCompDev> comps(ncomp);
  for (int c = 0; c < ncomp; ++c) {
    const sosgpu_aer_component &s = comp[c];
    if ((s.igranu != 1 && s.igranu != 2) || s.in > 0.0 || s.alpha0 > s.alphaf || !(s.wa > 0.0)) {
      ctx->err = "sosgpu_aerosols: component with igranu not 1 / 2, a positive imaginary index, alpha0 > alphaf or wa <= 0";
      return SOSGPU_ERR_ARG;
    }
    const auto key = std::make_tuple(s.rn, s.in, s.alpha0, s.alphaf);
    auto it = seen.find(key);
    if (it == seen.end()) {
      HostTable t{s.rn, s.in, s.alpha0, s.alphaf, alpha.size(), 0};
      if (!mie_grid(s.alpha0, s.alphaf, &alpha)) { ctx->err = "sosgpu_aerosols: size-parameter range outside CTE_MIE_DIM (SOS_MIE error 997)"; return SOSGPU_ERR_IER; }
      t.nrec = (int)(alpha.size() - t.rec0);
      it = seen.emplace(key, (int)tabs.size()).first;
      tabs.push_back(t);
    }
    const HostTable &t = tabs[it->second];
    comps[c] = GranuCompDev{(long long)t.rec0, t.nrec, s.igranu, s.alphaf, s.v1, s.v2, s.v3, s.wa};
  }
  CK(cudaSetDevice(ctx->device));
  cudaStream_t st = ctx->stream;
  const size_t nang = 2 * (size_t)nbmu + 1;
  SosFreeGuard guard(ctx);
  DevTables dt;
  int rc = alloc_tables(ctx, alpha.size(), nang, &dt, guard);
  if (rc != SOSGPU_OK) return rc;
  double *d_rmu = nullptr;
  CK(sos_dmalloc(ctx, &d_rmu, sizeof(double) * nang)); guard.add(d_rmu);
  CK(cudaMemcpyAsync(d_rmu, xmu, sizeof(double) * nang, cudaMemcpyHostToDevice, st));
  rc = launch_mie(ctx, nbmu, d_rmu, tabs, alpha, dt, guard);
  if (rc != SOSGPU_OK) return rc;
  return run_granu_models(ctx, nbmu, xmu, xhr, dt, comps, nullptr, false, mods, os_nb, comp_k, comp_phase, comp_ier, scal, coef, phase, model_ier, guard,
                          false);
}

// ---- the aerosol result file (SOS_AEROSOLS.F:2810-2832, formats 39-50) --------------------------
namespace {
void fmt_e(char *out, size_t cap, double x, int width, int dec)      // Fortran Ew.d (two-digit exponent; three digits drop the E)
{
  char m[64], s[64];
  if (x == 0.0) snprintf(s, sizeof s, "0.%0*dE+00", dec, 0);
  else {
    snprintf(m, sizeof m, "%.*E", dec - 1, fabs(x));
    char *e = strchr(m, 'E');
    const int ex = atoi(e + 1) + 1;
    *e = 0;
    char dig[32]; int nd = 0;
    for (char *c = m; *c; ++c) if (*c != '.') dig[nd++] = *c;
    dig[nd] = 0;
    if (ex <= -100 || ex >= 100) snprintf(s, sizeof s, "%s0.%s%c%03d", x < 0 ? "-" : "", dig, ex < 0 ? '-' : '+', abs(ex));
    else snprintf(s, sizeof s, "%s0.%sE%c%02d", x < 0 ? "-" : "", dig, ex < 0 ? '-' : '+', abs(ex));
  }
  if ((int)strlen(s) > width && s[0] == '0') memmove(s, s + 1, strlen(s));           // the optional leading zero goes first
  else if ((int)strlen(s) > width && s[0] == '-' && s[1] == '0') memmove(s + 1, s + 2, strlen(s + 1));
  if ((int)strlen(s) > width) { memset(s, '*', width); s[width] = 0; }
  snprintf(out, cap, "%*s", width, s);
}
void fmt_f(char *out, size_t cap, double x, int width, int dec)      // Fortran Fw.d
{
  char s[64];
  snprintf(s, sizeof s, "%.*f", dec, x);
  if ((int)strlen(s) > width && !strncmp(s, "0.", 2)) memmove(s, s + 1, strlen(s));
  else if ((int)strlen(s) > width && !strncmp(s, "-0.", 3)) memmove(s + 1, s + 2, strlen(s + 1));
  if ((int)strlen(s) > width) { memset(s, '*', width); s[width] = 0; }
  snprintf(out, cap, "%*s", width, s);
}
}  // namespace

extern "C" int sosgpu_write_aerosols(const char *path, int os_nb, double kmat1, double kmat2, double asym, double coef_tronca, double piztr,
                                     const double *alp, const double *beta11, const double *gamma12, const double *zeta)
{
  if (!path || os_nb < 0 || !alp || !beta11 || !gamma12 || !zeta) return SOSGPU_ERR_ARG;
  FILE *f = fopen(path, "w");
  if (!f) return SOSGPU_ERR_IER;
  char a[32], b[32], c[32], d[32];
  fmt_e(a, sizeof a, kmat1, 13, 5);  fprintf(f, "EXTINCTION CROSS SECTION (mic^2)     :%s\n", a);
  fmt_e(a, sizeof a, kmat2, 13, 5);  fprintf(f, "SCATTERING CROSS SECTION (mic^2)     :%s\n", a);
  fmt_e(a, sizeof a, asym, 13, 5);   fprintf(f, "ASYMMETRY FACTOR (no truncation)     :%s\n", a);
  fmt_f(a, sizeof a, coef_tronca, 9, 5); fprintf(f, "TRUNCATION COEFFICIENT               :%s\n", a);
  fmt_f(a, sizeof a, piztr, 9, 5);   fprintf(f, "SINGLE SCATTERING ALBEDO (truncation):%s\n", a);
  fprintf(f, "---------------------------------\n");
  fprintf(f, "PHASE MATRIX COEFFICIENTS FOR K=0 TO%4d\n", os_nb);
  fprintf(f, "ALPHA(K)        BETA11(K)       GAMMA12(K)      ZETA(K)\n");
  for (int k = 0; k <= os_nb; ++k) {
    fmt_e(a, sizeof a, alp[k], 15, 8); fmt_e(b, sizeof b, beta11[k], 15, 8); fmt_e(c, sizeof c, gamma12[k], 15, 8); fmt_e(d, sizeof d, zeta[k], 15, 8);
    fprintf(f, "%s %s %s %s\n", a, b, c, d);
  }
  const bool bad = ferror(f) != 0;
  if (fclose(f) != 0 || bad) return SOSGPU_ERR_IER;
  return SOSGPU_OK;
}

// ---- gfortran-ABI symbols (fixed strides of inc/SOS.h: angle vectors (-100:100), coefficients (0:200)) ----
namespace {
bool put_rec(FILE *f, const void *p, int n)
{
  return fwrite(&n, 4, 1, f) == 1 && fwrite(p, 1, (size_t)n, f) == (size_t)n && fwrite(&n, 4, 1, f) == 1;
}
}  // namespace

// SOS_MIE.F:205-206 (hidden lengths of FICMIE, FICLOG last).  Writes the unformatted MIE file; the trace file is not written.
extern "C" void sos_mie_(const int *mie_nbmu, const double *rmu, const double *chr, const double *rn, const double *in, const double *alphao,
                         const double *alphaf, const char *ficmie, const char *ficlog, int *ier, size_t len_ficmie, size_t len_ficlog)
{
  (void)chr; (void)ficlog; (void)len_ficlog;
  *ier = 0;
  const int n = *mie_nbmu;
  if (n < 1 || n > AC_MIE_NBMU_MAX) { *ier = -1; return; }
  if (trunc(*alphaf + *alphaf + 20) > AC_MIE_DIM) { printf(" Valeur AlphaMax trop grande devant CTE_MIE_DIM\n"); *ier = -1; return; }
  sosgpu_ctx *ctx = shim_ctx();
  if (!ctx) { printf("  SOS_MIE : no usable CUDA device\n"); *ier = -1; return; }
  const int cap = sosgpu_mie_count(*alphao, *alphaf);
  if (cap < 1) { *ier = -1; return; }
  const size_t nang = 2 * (size_t)n + 1;
  std::vector<float> rec(3 * (size_t)cap), im(cap * nang), qm(cap * nang), um(cap * nang);
  std::vector<double> g(cap);
  int nrec = 0;
  if (sosgpu_mie(ctx, n, rmu + (AC_MIE_NBMU_MAX - n), *rn, *in, *alphao, *alphaf, cap, rec.data(), g.data(), im.data(), qm.data(), um.data(),
                 &nrec) != SOSGPU_OK) { printf("  SOS_MIE : %s\n", sosgpu_last_error(ctx)); *ier = -1; return; }
  FILE *f = fopen(fstr(ficmie, len_ficmie).c_str(), "wb");
  if (!f) { printf(" Erreur a l'ouverture du fichier MIE\n"); *ier = -1; return; }
  char head[28];
  memcpy(head, rn, 8); memcpy(head + 8, in, 8); memcpy(head + 16, alphaf, 8); memcpy(head + 24, &n, 4);
  bool ok = put_rec(f, head, 28);
  std::vector<char> line(12 + 8 + 12 * nang);
  for (int k = 0; ok && k < nrec; ++k) {
    memcpy(line.data(), &rec[3 * (size_t)k], 12);
    memcpy(line.data() + 12, &g[k], 8);
    memcpy(line.data() + 20, &im[k * nang], 4 * nang);
    memcpy(line.data() + 20 + 4 * nang, &qm[k * nang], 4 * nang);
    memcpy(line.data() + 20 + 8 * nang, &um[k * nang], 4 * nang);
    ok = put_rec(f, line.data(), (int)line.size());
  }
  if (fclose(f) != 0 || !ok) { printf(" Erreur d'ecriture sur le fichier MIE\n"); *ier = -1; }
}

// SOS_AEROSOLS.F:4392-4394 (hidden length of FICMIE last).  Reads the MIE file; the trace block is not written.
extern "C" void sos_granu_(const char *ficmie, const int *igranu, const double *v1, const double *v2, const double *v3, const double *wa,
                           const int *mie_nbmu, const double *xmu, const int *trace, double *kmat1, double *kmat2, double *somme_nr,
                           double *p11, double *p12, double *p33, int *ier, size_t len_ficmie)
{
  (void)xmu; (void)trace;
  *ier = 0;
  const int n = *mie_nbmu;
  const size_t nang = 2 * (size_t)n + 1;
  FILE *f = fopen(fstr(ficmie, len_ficmie).c_str(), "rb");
  if (!f) { printf(" Erreur a l'ouverture du fichier MIE\n"); *ier = -1; return; }
  struct Closer { FILE *f; ~Closer() { fclose(f); } } closer{f};
  int l0 = 0, l1 = 0, nb = 0;
  double head[3];
  if (fread(&l0, 4, 1, f) != 1 || l0 != 28 || fread(head, 8, 3, f) != 3 || fread(&nb, 4, 1, f) != 1 || fread(&l1, 4, 1, f) != 1 || l1 != 28) {
    printf(" Erreur de lecture de la premiere ligne du fichier MIE\n"); *ier = -1; return;
  }
  if (n < 1 || n > AC_MIE_NBMU_MAX || nb != n) { printf(" Incoherence sur le nombre d'angles de Gauss\n NBMUMIE=%d MIE_NBMU=%d\n", nb, n); *ier = -1; return; }
  const int reclen = (int)(12 + 8 + 12 * nang);
  std::vector<float> rec, im, qm, um;
  std::vector<char> line(reclen);
  for (;;) {
    if (fread(&l0, 4, 1, f) != 1) break;
    if (l0 != reclen || fread(line.data(), 1, reclen, f) != (size_t)reclen || fread(&l1, 4, 1, f) != 1 || l1 != reclen) {
      printf(" Erreur de lecture sur le fichier MIE\n"); *ier = -1; return;
    }
    const float *p = (const float *)line.data();
    rec.insert(rec.end(), p, p + 3);
    const float *q = (const float *)(line.data() + 20);
    im.insert(im.end(), q, q + nang); qm.insert(qm.end(), q + nang, q + 2 * nang); um.insert(um.end(), q + 2 * nang, q + 3 * nang);
  }
  const int nrec = (int)(rec.size() / 3);
  sosgpu_ctx *ctx = shim_ctx();
  if (!ctx) { printf("  SOS_GRANU : no usable CUDA device\n"); *ier = -1; return; }
  std::vector<double> a(nang), b(nang), c(nang);
  double k[3] = {0, 0, 0};
  int e = -1;
  if (nrec < 1 || sosgpu_granu(ctx, n, nrec, rec.data(), im.data(), qm.data(), um.data(), head[2], *igranu, *v1, *v2, *v3, *wa, k, a.data(),
                               b.data(), c.data(), &e) != SOSGPU_OK || e != 0) {
    printf(" Erreur de lecture sur le fichier MIE\n  --> Probable tentative de lecture apres fin fichier\n"); *ier = -1; return;
  }
  *kmat1 = k[0]; *kmat2 = k[1]; *somme_nr = k[2];
  const size_t off = AC_MIE_NBMU_MAX - n;
  memcpy(p11 + off, a.data(), sizeof(double) * nang); memcpy(p12 + off, b.data(), sizeof(double) * nang); memcpy(p33 + off, c.data(), sizeof(double) * nang);
}

// SOS_AEROSOLS.F:3924-3928.  The trace lines are not written.
extern "C" void sos_decompo_legendre_(int *itronc, const int *trace, const int *mie_nbmu, const double *xmu, const double *xhr, const int *os_nb,
                                      double *p11, double *ttt, const double *p12, const double *p22, const double *p33, double *coef_tronca,
                                      double *z1, double *alp, double *beta11, double *beta22, double *gamma12, double *delta33, double *zeta,
                                      int *ier)
{
  (void)trace;
  const int n = *mie_nbmu;
  sosgpu_ctx *ctx = shim_ctx();
  if (!ctx) { printf("  SOS_DECOMPO_LEGENDRE : no usable CUDA device\n"); *ier = -1; return; }
  if (n < 1 || n > AC_MIE_NBMU_MAX) { *ier = -1; return; }
  const size_t off = AC_MIE_NBMU_MAX - n;
  int e = -1;
  if (sosgpu_decompo_legendre(ctx, itronc, n, xmu + off, xhr + off, *os_nb, p11 + off, ttt + off, p12 + off, p22 + off, p33 + off, coef_tronca, z1,
                              alp, beta11, beta22, gamma12, delta33, zeta, &e) != SOSGPU_OK || e != 0) {
    printf("  SOS_DECOMPO_LEGENDRE : %s\n", sosgpu_last_error(ctx));
    *ier = -1;
  }
}
