// sosgpu_api.cu -- host side of libsosgpu.so: C ABI (include/sosgpu.h), batch orchestration,
// Fourier-order waves, device memory.  No CPU compute fallback: without a CUDA device every
// compute entry returns SOSGPU_ERR_NO_DEVICE.
#include "sosgpu_host.h"

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <chrono>
#include <map>
#include <string>
#include <vector>

static inline int roundup(int x, int m) { return (x + m - 1) / m * m; }

// host-side arena that is uploaded in one copy; offsets are turned into device pointers afterwards.  With a pinned backing
// store (the context's staging buffer, sized up front) the upload is one true asynchronous DMA instead of a staged copy
// of pageable memory, and nothing is reallocated while the arena is filled.
struct Arena {
  std::vector<char> own;                                        // fallback storage (small single-routine calls)
  char *base = nullptr; size_t cap = 0, used = 0;
  Arena() {}
  Arena(char *pinned, size_t capacity) : base(pinned), cap(capacity) {}
  size_t put(const void *p, size_t bytes)
  {
    const size_t off = (used + 15) / 16 * 16;
    if (!base || off + bytes > cap) {                            // grow the private vector (or switch to it)
      if (base) { own.assign(base, base + used); base = nullptr; cap = 0; }
      own.resize(off + bytes);
    }
    char *dst = (base ? base : own.data()) + off;
    if (off > used) memset((base ? base : own.data()) + used, 0, off - used);
    if (p) memcpy(dst, p, bytes); else memset(dst, 0, bytes);
    used = off + bytes;
    return off;
  }
  template <class T> size_t putv(const std::vector<T> &v) { return put(v.data(), v.size() * sizeof(T)); }
  const char *data() const { return base ? base : own.data(); }
  size_t size() const { return used; }
};

// ---------------------------------------------------------------------------------------------
extern "C" int sosgpu_device_count(void)
{
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
  return n;
}

extern "C" void sosgpu_destroy(sosgpu_ctx *ctx);
extern "C" int sosgpu_create(sosgpu_ctx **out, int device)
{
  if (!out) return SOSGPU_ERR_ARG;
  *out = nullptr;
  if (sosgpu_device_count() <= 0) return SOSGPU_ERR_NO_DEVICE;
  sosgpu_ctx *ctx = new sosgpu_ctx();
  ctx->device = device;
  if (cudaSetDevice(device) != cudaSuccess || cudaStreamCreate(&ctx->stream) != cudaSuccess) {
    delete ctx;
    return SOSGPU_ERR_CUDA;
  }
  // Device memory of the batches comes from a stream-ordered pool that never trims: uploading and freeing a batch
  // per step (the host-buffer path) then costs no cudaMalloc/cudaFree (10-70 ms per step measured with plain calls).
  cudaMemPoolProps props = {};
  props.allocType = cudaMemAllocationTypePinned;
  props.handleTypes = cudaMemHandleTypeNone;
  props.location.type = cudaMemLocationTypeDevice;
  props.location.id = device;
  unsigned long long keep_all = ~0ull;
  if (cudaMemPoolCreate(&ctx->pool, &props) != cudaSuccess ||
      cudaMemPoolSetAttribute(ctx->pool, cudaMemPoolAttrReleaseThreshold, &keep_all) != cudaSuccess ||
      cudaMallocHost(&ctx->h_count, 16 * sizeof(int)) != cudaSuccess ||
      cudaEventCreate(&ctx->ev_a) != cudaSuccess || cudaEventCreate(&ctx->ev_b) != cudaSuccess ||
      cudaMalloc(&ctx->d_work_counter, 64) != cudaSuccess ||
      cudaDeviceGetAttribute(&ctx->num_sms, cudaDevAttrMultiProcessorCount, device) != cudaSuccess) {
    sosgpu_destroy(ctx);                                         // releases whatever was created so far
    return SOSGPU_ERR_CUDA;
  }
  for (int k = 0; k < 8; ++k)
    if (cudaEventCreateWithFlags(&ctx->ev_cnt[k], cudaEventDisableTiming) != cudaSuccess) { sosgpu_destroy(ctx); return SOSGPU_ERR_CUDA; }
  ctx->trace = getenv("SOS_TRACE") != nullptr;
  *out = ctx;
  return SOSGPU_OK;
}

extern "C" void sosgpu_destroy(sosgpu_ctx *ctx)
{
  if (!ctx) return;
  cudaSetDevice(ctx->device);
  if (ctx->stream) cudaStreamSynchronize(ctx->stream);
  cudaFree(ctx->grec_cache); cudaFree(ctx->cache_field); cudaFree(ctx->cache_kpool); cudaFree(ctx->d_work_counter); cudaFree(ctx->d_gather);
  sosgpu_comm_destroy(ctx);
  if (ctx->pool) cudaMemPoolDestroy(ctx->pool);
  if (ctx->h_count) cudaFreeHost(ctx->h_count);
  if (ctx->h_arena) cudaFreeHost(ctx->h_arena);
  if (ctx->ev_a) cudaEventDestroy(ctx->ev_a);
  if (ctx->ev_b) cudaEventDestroy(ctx->ev_b);
  for (int k = 0; k < 8; ++k) if (ctx->ev_cnt[k]) cudaEventDestroy(ctx->ev_cnt[k]);
  if (ctx->stream) cudaStreamDestroy(ctx->stream);
  delete ctx;
}

extern "C" const char *sosgpu_last_error(const sosgpu_ctx *ctx) { return ctx ? ctx->err.c_str() : "no context"; }
extern "C" long long sosgpu_launch_count(const sosgpu_ctx *ctx) { return ctx ? ctx->launches : 0; }
extern "C" double sosgpu_last_kernel_ms(const sosgpu_ctx *ctx) { return ctx ? (double)ctx->last_kernel_ms : 0.0; }
extern "C" int sosgpu_set_direct_models(sosgpu_ctx *ctx, const sosgpu_direct_models *dm)
{
  if (!ctx) return SOSGPU_ERR_NO_DEVICE;
  ctx->dm = dm ? *dm : sosgpu_direct_models{};
  return SOSGPU_OK;
}
extern "C" int sosgpu_set_options(sosgpu_ctx *ctx, size_t field_budget_bytes, int max_wave_orders)
{
  if (!ctx) return SOSGPU_ERR_ARG;
  if (field_budget_bytes) ctx->field_budget = field_budget_bytes;
  ctx->max_wave_orders = max_wave_orders;
  return SOSGPU_OK;
}

// ---------------------------------------------------------------------------------------------
// host preparation = the cheap per-call scalar work of SOS.F:496-586 and SOS_OS.F:678-839
static int prep_optics(const sosgpu_optics &o, HostOptics &h, bool os_level)
{
  if (o.nbmu < 1 || o.nbmu > SOSGPU_NBMU_MAX || o.os_nb < 2 || o.os_nb > SOSGPU_NB_MAX) return SOSGPU_ERR_ARG;
  h.N = o.nbmu; h.W = 2 * o.nbmu + 1; h.HB = roundup(3 * o.nbmu, 16); h.KP = 2 * h.HB;
  h.os_nb = o.os_nb; h.n0 = o.n0; h.imat_surf = o.imat_surf; h.ifresnel = o.ifresnel; h.ipolar = o.ipolar;
  h.igmax = o.igmax; h.ro = o.rho; h.ron = o.ron; h.ind_surf = o.ind_surf; h.zout = o.zout;
  h.a_trunc = o.a_trunc; h.piz = o.piz; h.piztr = o.piztr; h.surf = o.surf; h.n_surf_rec = o.n_surf_rec;
  (void)os_level;
  const int N = h.N, W = h.W;
  h.rmu.assign(o.rmu, o.rmu + W);
  h.ga.assign(o.ga, o.ga + W);
  h.alpha.assign(o.alpha, o.alpha + o.os_nb + 1);
  h.beta.assign(o.beta, o.beta + o.os_nb + 1);
  h.gamma.assign(o.gamma, o.gamma + o.os_nb + 1);
  h.zeta.assign(o.zeta, o.zeta + o.os_nb + 1);
  double aaa = o.ron / (2 - o.ron);                            // SOS_OS.F:678-684
  aaa = (1 - aaa) / (1 + 2 * aaa);
  h.beta2 = 0.5 * aaa;
  h.gamma2 = -aaa * std::sqrt(1.5);
  h.alpha2 = 3.0 * aaa;
  if (o.ipolar == 0) {                                         // SOS_OS.F:689-699
    h.gamma2 = 0.0; h.alpha2 = 0.0;
    for (int k = 0; k <= o.os_nb; ++k) { h.alpha[k] = 0.0; h.gamma[k] = 0.0; h.zeta[k] = 0.0; }
  }
  if (o.n0 > 0) {                                              // SOS_OS.F:706-715
    if (o.n0 > N) return SOSGPU_ERR_ARG;
    h.tab = -h.rmu[o.n0 + N];
  } else {
    h.tab = -std::cos(std::acos(-1.0) * o.tetas / 180.0);
  }
  h.rmu[N] = h.tab;
  h.limb = (h.tab == 0.0);
  if (o.imat_surf == 1 && (o.surf == nullptr || o.n0 <= 0)) return SOSGPU_ERR_ARG;
  h.f11.assign(N + 1, 0.0); h.f12.assign(N + 1, 0.0); h.f33.assign(N + 1, 0.0);
  if (o.ifresnel == 1) {                                       // SOS_MAT_FRESNEL_PLAN_REFL, SOS_OS.F:1753-1780
    for (int j = 0; j <= N; ++j) {
      const double mu = (j == 0) ? -h.rmu[N] : h.rmu[j + N];
      const double ind2 = o.ind_surf * o.ind_surf, mu2 = mu * mu;
      const double x = std::sqrt(ind2 - 1.0 + mu2);
      const double rl = (ind2 * mu - x) / (ind2 * mu + x), rr = (mu - x) / (mu + x);
      h.f11[j] = (rl * rl + rr * rr) / 2.0;
      if (o.ipolar == 1) { h.f12[j] = (rl * rl - rr * rr) / 2.0; h.f33[j] = rl * rr; }
    }
  }
  h.f11sun = h.f11[0]; h.f12sun = h.f12[0];
  return SOSGPU_OK;
}

static int prep_term(const sosgpu_term &t, const HostOptics &o, HostTerm &h, bool os_level)
{
  if (t.nt < 1 || t.nt > SOSGPU_NT_MAX) return SOSGPU_ERR_ARG;
  const int nt = t.nt;
  h.optics = t.optics; h.group = t.group; h.nt = nt; h.LP = roundup(nt + 1, 8); h.aik = t.aik; h.ier = 0;
  h.h.assign(t.h, t.h + nt + 1);
  h.xdel.assign(t.pcaer, t.pcaer + nt + 1);
  h.ydel.assign(t.pcmol, t.pcmol + nt + 1);
  h.ttot_vrai = h.h[nt];                                       // SOS.F:518
  bool lta = true;
  if (!os_level) {
    std::vector<double> htr(nt + 1, 0.0);
    htr[0] = h.h[0];
    if (o.a_trunc != 0.0) {                                    // SOS.F:525-537
      for (int i = 1; i <= nt; ++i) {
        const double dh = h.h[i] - h.h[i - 1];
        const double va = h.xdel[i] * dh;
        const double vatr = va * (1 - o.piz * 0.5 * o.a_trunc);
        const double vr = h.ydel[i] * dh;
        const double vg = (1 - h.xdel[i] - h.ydel[i]) * dh;
        htr[i] = (vatr + vr + vg) + htr[i - 1];
        h.xdel[i] = vatr / (vatr + vr + vg);
        h.ydel[i] = vr / (vatr + vr + vg);
      }
    }
    for (int i = 0; i <= nt; ++i) {                            // SOS.F:539-543
      if (o.a_trunc != 0.0) h.h[i] = htr[i];
      h.xdel[i] = h.xdel[i] * o.piztr;
      if (h.xdel[i] != 0.0) lta = false;
    }
    h.iborm = lta ? 2 : o.os_nb;                               // SOS.F:549-550
  } else {
    h.iborm = o.os_nb;
  }
  h.ttot_tronc = h.h[nt];                                      // SOS.F:586
  h.jout = -1; h.zz = 0.0;
  if (o.zout == -1) h.tauout = h.h[0];                         // SOS.F:567-583
  else {
    if ((o.zout < 0) || (o.zout > 120.0)) { h.ier = -1; h.tauout = 0.0; }   // SOS_OS.F:811
    else {
      int j = 1;
      while (j < nt && o.zout < t.zprof[j]) ++j;
      h.jout = j;
      h.zz = (o.zout - t.zprof[j - 1]) / (t.zprof[j] - t.zprof[j - 1]);
      h.tauout = (1 - h.zz) * h.h[j - 1] + h.zz * h.h[j];
    }
  }
  h.dt.assign(nt + 2, 0.0); h.inv.assign(nt + 2, 0.0);           // +2: TMA copies of the tables are 16-byte granular
  h.ch.assign(nt + 2, 0.0); h.cf.assign(nt + 2, 0.0);            // +1: TMA table copies are 16-byte granular
  for (int i = 0; i < nt; ++i) { h.dt[i] = h.h[i + 1] - h.h[i]; h.inv[i] = 1.0 / h.dt[i]; }
  h.eground = std::exp(h.h[nt] / o.tab);                       // CH / CF per level are filled on the device (k_beam)
  return SOSGPU_OK;
}

static void free_batch_device(sosgpu_ctx *ctx, sosgpu_batch *b)
{
  const bool tr = getenv("SOS_TRACE") != nullptr;
  auto now = [] { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
  const double t0 = now();
  if (ctx && b->d_field && b->field_bytes > ctx->cache_field_bytes) {
    cudaFree(ctx->cache_field); ctx->cache_field = b->d_field; ctx->cache_field_bytes = b->field_bytes;
  } else cudaFree(b->d_field);
  if (ctx && b->d_kpool && b->kpool_bytes > ctx->cache_kpool_bytes) {
    cudaFree(ctx->cache_kpool); ctx->cache_kpool = b->d_kpool; ctx->cache_kpool_bytes = b->kpool_bytes;
  } else cudaFree(b->d_kpool);
  const double t1 = now();
  sos_dfree(ctx, b->d_arena); sos_dfree(ctx, b->d_optics); sos_dfree(ctx, b->d_terms); sos_dfree(ctx, b->d_att); sos_dfree(ctx, b->d_i4);
  sos_dfree(ctx, b->d_rec); sos_dfree(ctx, b->d_emoins); sos_dfree(ctx, b->d_eplus);
  sos_dfree(ctx, b->d_nf); sos_dfree(ctx, b->d_nsc); sos_dfree(ctx, b->d_rsn); sos_dfree(ctx, b->d_done); sos_dfree(ctx, b->d_gnrec);
  sos_dfree(ctx, b->d_group_start); sos_dfree(ctx, b->d_group_terms);
  sos_dfree(ctx, b->d_items); sos_dfree(ctx, b->d_ksets); sos_dfree(ctx, b->d_item_of);
  sos_dfree(ctx, b->d_list[0]); sos_dfree(ctx, b->d_list[1]); sos_dfree(ctx, b->d_count);
  sos_dfree(ctx, b->d_tg); sos_dfree(ctx, b->d_tphi); sos_dfree(ctx, b->d_tout);
  const double t2 = now();
  if (ctx && b->d_grec && b->grec_bytes > ctx->grec_cache_bytes) {
    cudaFree(ctx->grec_cache);
    ctx->grec_cache = b->d_grec; ctx->grec_cache_bytes = b->grec_bytes;
  } else cudaFree(b->d_grec);
  const double t3 = now();
  if (b->ev0) cudaEventDestroy(b->ev0);
  if (b->ev1) cudaEventDestroy(b->ev1);
  if (b->evt0) cudaEventDestroy(b->evt0);
  if (b->evt1) cudaEventDestroy(b->evt1);
  for (cudaEvent_t e : b->ev_order) cudaEventDestroy(e);
  b->ev_order.clear();
  if (tr) fprintf(stderr, "batch_free: pools %.2f ms, small %.2f ms, group buffer %.2f ms, events %.2f ms\n", t1 - t0, t2 - t1,
                  t3 - t2, now() - t3);
}

extern "C" void sosgpu_batch_free(sosgpu_ctx *ctx, sosgpu_batch *b)
{
  if (!b) return;
  if (ctx) cudaSetDevice(ctx->device);
  free_batch_device(ctx, b);
  delete b;
}

static int upload_impl(sosgpu_ctx *ctx, const sosgpu_optics *optics, int noptics, const sosgpu_term *terms,
                       int nterm, int ngroup, bool os_level, sosgpu_batch **out, const int *iborm = nullptr)
{
  if (!ctx) return SOSGPU_ERR_NO_DEVICE;
  if (!optics || !terms || noptics < 1 || nterm < 1 || ngroup < 1 || !out) { ctx->err = "bad arguments"; return SOSGPU_ERR_ARG; }
  CK(cudaSetDevice(ctx->device));
  const auto t_up0 = std::chrono::steady_clock::now();
  sosgpu_batch *b = new sosgpu_batch();
  struct Guard {                                                // frees a half-built batch on every error return
    sosgpu_ctx *c; sosgpu_batch *b;
    ~Guard() { if (b) { free_batch_device(c, b); delete b; } }
  } guard{ctx, b};
  b->nterm = nterm; b->noptics = noptics; b->ngroup = ngroup;
  b->ho.resize(noptics); b->ht.resize(nterm);
  for (int i = 0; i < noptics; ++i) {
    int rc = prep_optics(optics[i], b->ho[i], os_level);
    if (rc != SOSGPU_OK) { ctx->err = "invalid optics entry"; return rc; }
    b->maxHB = std::max(b->maxHB, b->ho[i].HB); b->maxW = std::max(b->maxW, b->ho[i].W);
    b->maxKP = std::max(b->maxKP, b->ho[i].KP); b->maxNB = std::max(b->maxNB, b->ho[i].os_nb);
    if (b->ho[i].ifresnel == 1) b->any_fresnel = 1;
  }
  int max_att = 0;
  for (int i = 0; i < nterm; ++i) {
    if (terms[i].optics < 0 || terms[i].optics >= noptics || terms[i].group < 0 || terms[i].group >= ngroup) {
      ctx->err = "term refers to a missing optics / group"; return SOSGPU_ERR_ARG;
    }
    int rc = prep_term(terms[i], b->ho[terms[i].optics], b->ht[i], os_level);
    if (rc != SOSGPU_OK) { ctx->err = "invalid term entry"; return rc; }
    if (os_level && iborm) b->ht[i].iborm = std::min(std::max(iborm[i], 0), b->ho[terms[i].optics].os_nb);
    b->smax = std::max(b->smax, b->ht[i].iborm + 1);
    max_att = std::max(max_att, b->ht[i].nt * b->ho[terms[i].optics].N);
  }
  b->rs_dev = b->maxNB + 1;
  b->w_dev = b->maxW;

  const auto t_up1 = std::chrono::steady_clock::now();
  // ---- constant arena ----
  // size of the arena: per optics the angle / coefficient vectors (+ surface records), per term seven level arrays
  size_t need_bytes = 4096;
  for (int i = 0; i < noptics; ++i) {
    const HostOptics &h = b->ho[i];
    need_bytes += (size_t)(2 * h.W + 4 * (h.os_nb + 1) + 3 * (h.N + 1)) * 8 + 10 * 16;
    if (h.imat_surf == 1) need_bytes += (size_t)std::min(h.n_surf_rec, h.os_nb + 1) * 9 * h.N * h.N * sizeof(float) + 16;
  }
  for (int i = 0; i < nterm; ++i) need_bytes += (size_t)7 * ((b->ht[i].nt + 2) * 8 + 16);
  if (need_bytes > ctx->h_arena_cap) {                           // pinned staging buffer of the context, grown on demand
    if (ctx->h_arena) cudaFreeHost(ctx->h_arena);
    ctx->h_arena = nullptr; ctx->h_arena_cap = 0;
    if (cudaMallocHost(&ctx->h_arena, need_bytes + need_bytes / 4) == cudaSuccess) ctx->h_arena_cap = need_bytes + need_bytes / 4;
    else cudaGetLastError();                                     // no pinned memory: the arena falls back to pageable storage
  }
  Arena ar(ctx->h_arena, ctx->h_arena_cap);
  b->optics_dev.resize(noptics);
  std::vector<size_t> o_off(noptics * 10);
  for (int i = 0; i < noptics; ++i) {
    HostOptics &h = b->ho[i];
    size_t *o = &o_off[i * 10];
    o[0] = ar.putv(h.rmu); o[1] = ar.putv(h.ga); o[2] = ar.putv(h.alpha); o[3] = ar.putv(h.beta);
    o[4] = ar.putv(h.gamma); o[5] = ar.putv(h.zeta); o[6] = ar.putv(h.f11); o[7] = ar.putv(h.f12); o[8] = ar.putv(h.f33);
    o[9] = 0;
    if (h.imat_surf == 1) {
      const size_t need = (size_t)std::min(h.n_surf_rec, h.os_nb + 1);
      if ((int)need < 1) { ctx->err = "surface matrix has no record"; return SOSGPU_ERR_ARG; }
      o[9] = ar.put(h.surf, need * 9 * h.N * h.N * sizeof(float));
    }
  }
  std::vector<size_t> t_off(nterm * 7);
  size_t att_total = 0, i4_total = 0;
  std::vector<size_t> att_off(nterm), i4_off(nterm), lvl_off(nterm);
  size_t lvl_total = 0;
  for (int i = 0; i < nterm; ++i) {
    HostTerm &h = b->ht[i];
    size_t *o = &t_off[i * 7];
    o[0] = ar.putv(h.h); o[1] = ar.putv(h.xdel); o[2] = ar.putv(h.ydel); o[3] = ar.putv(h.dt);
    o[4] = ar.putv(h.inv); o[5] = ar.putv(h.ch); o[6] = ar.putv(h.cf);
    // 16-byte aligned; rows >= nt up to the end of the last 64-level chunk (+2) stay zero: k_step2 reads them as a = 0
    att_off[i] = att_total; att_total += (((size_t)(roundup(h.nt + 1, SOS_CH) + 2) * b->ho[h.optics].N + 1) & ~(size_t)1);
    i4_off[i] = i4_total; i4_total += (size_t)12 * b->ho[h.optics].N;
    lvl_off[i] = lvl_total; lvl_total += (size_t)((h.nt + 2 + 1) & ~1);
  }
  b->i4_total = i4_total;
  CK(sos_dmalloc(ctx, &b->d_arena, ar.size()));
  CK(cudaMemcpyAsync(b->d_arena, ar.data(), ar.size(), cudaMemcpyHostToDevice, ctx->stream));
  // a, g, 1-a-g, pup, qup, pdn, qdn and the five first-order tables (k_att) + the two level tables of the order-1 source (k_beam)
  CK(sos_dmalloc(ctx, &b->d_att, (12 * att_total + 2 * lvl_total) * sizeof(double)));
  CK(cudaMemsetAsync(b->d_att, 0, (12 * att_total + 2 * lvl_total) * sizeof(double), ctx->stream));
  CK(sos_dmalloc(ctx, &b->d_i4, i4_total * sizeof(double)));
  for (int i = 0; i < noptics; ++i) {
    HostOptics &h = b->ho[i];
    OpticsDev &d = b->optics_dev[i];
    size_t *o = &o_off[i * 10];
    d.nbmu = h.N; d.W = h.W; d.HB = h.HB; d.KP = h.KP; d.os_nb = h.os_nb; d.n0 = h.n0; d.imat_surf = h.imat_surf;
    d.ifresnel = h.ifresnel; d.ipolar = h.ipolar; d.igmax = h.igmax; d.tab = h.tab; d.ro = h.ro; d.ron = h.ron;
    d.ind_surf = h.ind_surf; d.zout = h.zout; d.beta2 = h.beta2; d.gamma2 = h.gamma2; d.alpha2 = h.alpha2;
    d.f11sun = h.f11sun; d.f12sun = h.f12sun;
    d.rmu = (const double *)(b->d_arena + o[0]); d.ga = (const double *)(b->d_arena + o[1]);
    d.alpha = (const double *)(b->d_arena + o[2]); d.beta = (const double *)(b->d_arena + o[3]);
    d.gamma = (const double *)(b->d_arena + o[4]); d.zeta = (const double *)(b->d_arena + o[5]);
    d.f11 = (const double *)(b->d_arena + o[6]); d.f12 = (const double *)(b->d_arena + o[7]);
    d.f33 = (const double *)(b->d_arena + o[8]);
    d.surf = h.imat_surf == 1 ? (const float *)(b->d_arena + o[9]) : nullptr;
    d.n_surf_rec = h.n_surf_rec;
  }
  b->terms_dev.resize(nterm);
  for (int i = 0; i < nterm; ++i) {
    HostTerm &h = b->ht[i];
    TermDev &d = b->terms_dev[i];
    size_t *o = &t_off[i * 7];
    d.optics = h.optics; d.group = h.group; d.nt = h.nt; d.LP = h.LP; d.iborm = h.iborm; d.jout = h.jout; d.zz = h.zz;
    d.aik = h.aik; d.eground = h.eground;
    d.h = (const double *)(b->d_arena + o[0]); d.xdel = (const double *)(b->d_arena + o[1]);
    d.ydel = (const double *)(b->d_arena + o[2]); d.dt = (const double *)(b->d_arena + o[3]);
    d.inv = (const double *)(b->d_arena + o[4]); d.ch = (const double *)(b->d_arena + o[5]);
    d.cf = (const double *)(b->d_arena + o[6]);
    d.att = b->d_att + att_off[i];
    d.gco = b->d_att + att_total + att_off[i];
    d.bco = b->d_att + 2 * att_total + att_off[i];
    d.pup = b->d_att + 3 * att_total + att_off[i]; d.qup = b->d_att + 4 * att_total + att_off[i];
    d.pdn = b->d_att + 5 * att_total + att_off[i]; d.qdn = b->d_att + 6 * att_total + att_off[i];
    d.o1u1 = b->d_att + 7 * att_total + att_off[i]; d.o1u2 = b->d_att + 8 * att_total + att_off[i];
    d.adn = b->d_att + 9 * att_total + att_off[i];
    d.o1d1 = b->d_att + 10 * att_total + att_off[i]; d.o1d2 = b->d_att + 11 * att_total + att_off[i];
    d.sxd = b->d_att + 12 * att_total + lvl_off[i]; d.syd = b->d_att + 12 * att_total + lvl_total + lvl_off[i];
    d.i4 = b->d_i4 + i4_off[i];
  }
  CK(sos_dmalloc(ctx, &b->d_optics, noptics * sizeof(OpticsDev)));
  CK(sos_dmalloc(ctx, &b->d_terms, nterm * sizeof(TermDev)));
  CK(cudaMemcpyAsync(b->d_optics, b->optics_dev.data(), noptics * sizeof(OpticsDev), cudaMemcpyHostToDevice, ctx->stream));
  CK(cudaMemcpyAsync(b->d_terms, b->terms_dev.data(), nterm * sizeof(TermDev), cudaMemcpyHostToDevice, ctx->stream));

  // groups: CSR of term indices in original order (the reference aggregates in CKD loop order)
  b->group_start.assign(ngroup + 1, 0);
  for (int i = 0; i < nterm; ++i) b->group_start[b->ht[i].group + 1]++;
  for (int g = 0; g < ngroup; ++g) b->group_start[g + 1] += b->group_start[g];
  b->group_terms.resize(nterm);
  {
    std::vector<int> fill(b->group_start.begin(), b->group_start.end() - 1);
    for (int i = 0; i < nterm; ++i) b->group_terms[fill[b->ht[i].group]++] = i;
  }
  b->group_optics.assign(ngroup, -1);
  for (int g = 0; g < ngroup; ++g)
    if (b->group_start[g + 1] > b->group_start[g]) b->group_optics[g] = b->ht[b->group_terms[b->group_start[g]]].optics;
  CK(sos_dmalloc(ctx, &b->d_group_start, (ngroup + 1) * sizeof(int)));
  CK(sos_dmalloc(ctx, &b->d_group_terms, nterm * sizeof(int)));
  CK(cudaMemcpyAsync(b->d_group_start, b->group_start.data(), (ngroup + 1) * sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
  CK(cudaMemcpyAsync(b->d_group_terms, b->group_terms.data(), nterm * sizeof(int), cudaMemcpyHostToDevice, ctx->stream));

  const size_t per = (size_t)b->rs_dev * 3 * b->w_dev;
  CK(sos_dmalloc(ctx, &b->d_rec, (size_t)nterm * per * sizeof(double)));
  // The group sums are handed to NCCL (sosgpu_batch_group_buffer), so they stay a plain cudaMalloc block; a freed
  // batch parks it in the context because cudaFree of it was measured at 1.5-670 ms (device-wide synchronisation).
  // + per-group tail (8 scalars, rs_dev series-length indicators) so that ONE ncclReduce carries everything (sosgpu_comm.cu)
  b->grec_bytes = (size_t)ngroup * (per + 8 + b->rs_dev) * sizeof(double);
  if (ctx->grec_cache && ctx->grec_cache_bytes >= b->grec_bytes) {
    b->d_grec = ctx->grec_cache; b->grec_bytes = ctx->grec_cache_bytes;
    ctx->grec_cache = nullptr; ctx->grec_cache_bytes = 0;
  } else CK(cudaMalloc(&b->d_grec, b->grec_bytes));
  CK(sos_dmalloc(ctx, &b->d_emoins, nterm * sizeof(double)));
  CK(sos_dmalloc(ctx, &b->d_eplus, nterm * sizeof(double)));
  CK(sos_dmalloc(ctx, &b->d_nf, nterm * sizeof(int)));
  CK(sos_dmalloc(ctx, &b->d_nsc, (size_t)nterm * b->rs_dev * sizeof(int)));
  CK(sos_dmalloc(ctx, &b->d_rsn, (size_t)nterm * b->rs_dev * sizeof(int)));
  CK(sos_dmalloc(ctx, &b->d_done, nterm * sizeof(int)));
  CK(sos_dmalloc(ctx, &b->d_gnrec, ngroup * sizeof(int)));
  CK(sos_dmalloc(ctx, &b->d_count, 2 * sizeof(int)));
  b->h_count = ctx->h_count;
  CK(cudaEventCreate(&b->ev0)); CK(cudaEventCreate(&b->ev1));
  CK(cudaEventCreate(&b->evt0)); CK(cudaEventCreate(&b->evt1));

  sos_launch_att(b->d_terms, b->d_optics, nterm, max_att, ctx->stream);   // k_att + k_beam
  ctx->launches += 2;
  CK(cudaGetLastError());
  const auto t_up2 = std::chrono::steady_clock::now();
  CK(cudaStreamSynchronize(ctx->stream));
  if (getenv("SOS_TRACE")) {
    const auto t_up3 = std::chrono::steady_clock::now();
    auto ms = [](auto a, auto b) { return std::chrono::duration<double, std::milli>(b - a).count(); };
    fprintf(stderr, "batch_upload: host prep %.2f ms, arena + allocations + copies %.2f ms, wait %.2f ms\n", ms(t_up0, t_up1),
            ms(t_up1, t_up2), ms(t_up2, t_up3));
  }
  guard.b = nullptr;
  *out = b;
  return SOSGPU_OK;
}

extern "C" int sosgpu_batch_upload(sosgpu_ctx *ctx, const sosgpu_optics *optics, int noptics,
                                   const sosgpu_term *terms, int nterm, int ngroup, sosgpu_batch **batch)
{
  return upload_impl(ctx, optics, noptics, terms, nterm, ngroup, false, batch);
}

extern "C" int sosgpu_batch_upload_os(sosgpu_ctx *ctx, const sosgpu_optics *optics, int noptics,
                                      const sosgpu_term *terms, int nterm, int ngroup, const int *iborm,
                                      sosgpu_batch **batch)
{
  return upload_impl(ctx, optics, noptics, terms, nterm, ngroup, true, batch, iborm);
}

template <class T> static int ensure(sosgpu_ctx *ctx, T **p, size_t *cap, size_t need)
{
  if (need <= *cap) return SOSGPU_OK;
  if (*p) sos_dfree(ctx, *p);
  *p = nullptr; *cap = 0;
  CK(sos_dmalloc(ctx, p, need * sizeof(T)));
  *cap = need;
  return SOSGPU_OK;
}

struct WaveItemHost { int term, is; };

static int run_impl(sosgpu_ctx *ctx, sosgpu_batch *b, double *jdump_dev, int forced_first_n, const double *x_in_dev)
{
  (void)forced_first_n; (void)x_in_dev;
  CK(cudaSetDevice(ctx->device));
  cudaStream_t st = ctx->stream;
  const int nterm = b->nterm;
  const size_t per = (size_t)b->rs_dev * 3 * b->w_dev;
  b->stats = sosgpu_stats{};
  struct ItemFlops { int term, is; double flops; };
  std::vector<ItemFlops> item_flops;                             // for stats.useful_flops, once the Fourier counts are final
  b->reduced = false;
  CK(cudaEventRecord(b->evt0, st));
  CK(cudaMemsetAsync(b->d_rec, 0, (size_t)nterm * per * sizeof(double), st));
  CK(cudaMemsetAsync(b->d_i4, 0, b->i4_total * sizeof(double), st));
  CK(cudaMemsetAsync(b->d_done, 0, nterm * sizeof(int), st));
  CK(cudaMemsetAsync(b->d_nf, 0, nterm * sizeof(int), st));
  CK(cudaMemsetAsync(b->d_nsc, 0, (size_t)nterm * b->rs_dev * sizeof(int), st));
  CK(cudaMemsetAsync(b->d_rsn, 0xff, (size_t)nterm * b->rs_dev * sizeof(int), st));
  CK(cudaMemsetAsync(b->d_emoins, 0, nterm * sizeof(double), st));
  CK(cudaMemsetAsync(b->d_eplus, 0, nterm * sizeof(double), st));

  std::vector<int> done(nterm, 0);
  for (int i = 0; i < nterm; ++i) if (b->ht[i].ier != 0 || b->ho[b->ht[i].optics].limb) done[i] = 1;
  {
    bool any = false;
    for (int i = 0; i < nterm; ++i) any |= (done[i] != 0);
    if (any) CK(cudaMemcpyAsync(b->d_done, done.data(), nterm * sizeof(int), cudaMemcpyHostToDevice, st));
  }

  // field budget: the configured cap, but never more than 70 % of what the device can still give plus what this batch /
  // context already holds: a large batch on a smaller or shared GPU gets narrower waves, not an out-of-memory error
  size_t budget = ctx->field_budget;
  if (b->field_bytes == 0 && ctx->cache_field_bytes == 0) {      // only when a pool has to be allocated (the query is not free)
    size_t free_b = 0, total_b = 0;
    if (cudaMemGetInfo(&free_b, &total_b) == cudaSuccess) budget = std::min(budget, (size_t)(0.7 * (double)free_b));
  }
  int s0 = 0;
  std::vector<ItemDev> items;
  std::vector<KsetDev> ksets;
  std::vector<int> item_of;
  while (s0 < b->smax) {
    // ---- active terms and wave width ----
    std::vector<int> act;
    size_t bytes_per_order = 0;
    for (int i = 0; i < nterm; ++i)
      if (!done[i] && s0 <= b->ht[i].iborm) {
        act.push_back(i);
        bytes_per_order += (size_t)2 * SOS_XSIZE(b->ho[b->ht[i].optics].KP, b->ht[i].nt + 1) * sizeof(double);
      }
    if (act.empty()) break;
    int ws = b->smax - s0;
    const int by_par = std::max(8, (int)((1024 + act.size() - 1) / act.size()));
    ws = std::min(ws, by_par);
    if (ctx->max_wave_orders > 0) ws = std::min(ws, ctx->max_wave_orders);
    ws = std::min<size_t>(ws, std::max<size_t>(1, budget / std::max<size_t>(bytes_per_order, 1)));
    const int s1 = s0 + ws;

    // ---- kernel sets: unique (optics, s) ----
    std::map<std::pair<int, int>, int> kmap;
    ksets.clear(); items.clear();
    item_of.assign((size_t)nterm * ws, -1);
    size_t kbytes = 0, fbytes = 0, sbytes = 0;
    auto al = [](size_t x) { return (x + 255) / 256 * 256; };
    std::vector<size_t> koff, foff, soff;
    // Work units are handed out dynamically in list order: longest profiles first bounds the tail of every launch to the
    // shortest units; terms of one wavelength stay adjacent within a length class (they share the A operand in L2).
    std::vector<int> act_sorted(act);
    std::stable_sort(act_sorted.begin(), act_sorted.end(), [&](int x, int y) {
      const int cx = (b->ht[x].nt + SOS_CH) / SOS_CH, cy = (b->ht[y].nt + SOS_CH) / SOS_CH;
      if (cx != cy) return cx > cy;
      return b->ht[x].optics < b->ht[y].optics;
    });
    for (int ti : act_sorted) {
      const HostTerm &ht = b->ht[ti];
      const HostOptics &ho = b->ho[ht.optics];
      for (int s = s0; s < s1 && s <= ht.iborm; ++s) {
        auto key = std::make_pair(ht.optics, s);
        auto itk = kmap.find(key);
        int kid;
        if (itk == kmap.end()) {
          kid = (int)ksets.size();
          kmap[key] = kid;
          KsetDev k{};
          k.optics = ht.optics; k.is = s; k.dual = (s <= 2) ? 1 : 0; k.beta0 = (s == 0) ? 1.0 : 0.0;
          ksets.push_back(k);
          koff.push_back(kbytes);
          const size_t W = ho.W, KP = ho.KP;
          kbytes += al(3 * (ho.os_nb + 2) * W * 8) + al(6 * W * W * 8) + al(3 * W * 8) + al(KP * KP * 8) * (k.dual ? 2 : 1) + 5 * al(KP * 8) + al(16 * KP * 8);
        } else kid = itk->second;
        ItemDev it{};
        it.term = ti; it.is = s; it.kset = kid; it.n = 1; it.active = 1; it.reason = -1;
        item_of[(size_t)ti * ws + (s - s0)] = (int)items.size();
        items.push_back(it);
        foff.push_back(fbytes);
        fbytes += al((size_t)2 * SOS_XSIZE(ho.KP, ht.nt + 1) * 8);
        soff.push_back(sbytes);
        const size_t N = ho.N;
        sbytes += al((3 * 6 * N + 3 * N) * 8) + (ht.jout >= 0 ? al((2 * 6 * N + 2 * 3 * N) * 8) : 0);
      }
    }
    const size_t nitem = items.size(), nk = ksets.size();
    // ---- pools ----
    // The two multi-GB wave pools are plain cudaMalloc blocks parked in the context between batches (growing the
    // stream-ordered pool by several GB costs ~0.5 s on first use; cudaFree synchronises the device).
    if (fbytes + sbytes > b->field_bytes && ctx->cache_field_bytes >= fbytes + sbytes) {
      if (b->d_field) cudaFree(b->d_field);
      b->d_field = ctx->cache_field; b->field_bytes = ctx->cache_field_bytes;
      ctx->cache_field = nullptr; ctx->cache_field_bytes = 0;
      CK(cudaMemsetAsync(b->d_field, 0, b->field_bytes, st));     // see below: pads must be finite
    }
    if (kbytes > b->kpool_bytes && ctx->cache_kpool_bytes >= kbytes) {
      if (b->d_kpool) cudaFree(b->d_kpool);
      b->d_kpool = ctx->cache_kpool; b->kpool_bytes = ctx->cache_kpool_bytes;
      ctx->cache_kpool = nullptr; ctx->cache_kpool_bytes = 0;
    }
    if (fbytes + sbytes > b->field_bytes) {
      if (b->d_field) cudaFree(b->d_field);
      b->d_field = nullptr; b->field_bytes = 0;
      CK(cudaMalloc(&b->d_field, fbytes + sbytes));
      b->field_bytes = fbytes + sbytes;
      // Zeroed once per allocation: pad rows / levels / columns of the fields only ever meet zero coefficients of the
      // packed operators, so they merely have to stay finite (no NaN bit patterns of fresh memory); every valid
      // element and every per-item state array is rewritten by the order-1 kernel and k_init of each wave.
      CK(cudaMemsetAsync(b->d_field, 0, fbytes + sbytes, st));
    }
    if (kbytes > b->kpool_bytes) {
      if (b->d_kpool) cudaFree(b->d_kpool);
      b->d_kpool = nullptr; b->kpool_bytes = 0;
      CK(cudaMalloc(&b->d_kpool, kbytes));
      b->kpool_bytes = kbytes;
    }
    CK(cudaMemsetAsync(b->d_kpool, 0, kbytes, st));              // PSL/RSL/TSL rely on zero-initialised storage
    for (size_t i = 0; i < nk; ++i) {
      KsetDev &k = ksets[i];
      const HostOptics &ho = b->ho[k.optics];
      const size_t W = ho.W, KP = ho.KP;
      char *p = b->d_kpool + koff[i];
      k.basis = (double *)p; p += al(3 * (ho.os_nb + 2) * W * 8);
      k.ker = (double *)p;   p += al(6 * W * W * 8);
      k.xpl = (double *)p;   p += al(3 * W * 8);
      k.apackA = (double *)p; p += al(KP * KP * 8);
      if (k.dual) { k.apackR = (double *)p; p += al(KP * KP * 8); }
      k.c1 = (double *)p; p += al(KP * 8);
      k.c2 = (double *)p; p += al(KP * 8);
      k.fz1 = (double *)p; p += al(KP * 8);
      k.fz2 = (double *)p; p += al(KP * 8);
      k.urow = (double *)p; p += al(KP * 8);
      k.vpack = (double *)p; p += al(16 * KP * 8);
    }
    for (size_t i = 0; i < nitem; ++i) {
      ItemDev &it = items[i];
      const HostTerm &ht = b->ht[it.term];
      const HostOptics &ho = b->ho[ht.optics];
      const size_t N = ho.N;
      char *p = b->d_field + foff[i];
      it.x[0] = (double *)p;
      it.x[1] = it.x[0] + SOS_XSIZE(ho.KP, ht.nt + 1);
      char *q = b->d_field + fbytes + soff[i];
      it.hist_a = (double *)q; it.hist_d = it.hist_a + 6 * N; it.sum3 = it.hist_d + 6 * N; it.rii = it.sum3 + 6 * N;
      if (ht.jout >= 0) {
        q += al((3 * 6 * N + 3 * N) * 8);
        it.sumout = (double *)q; it.riiout = it.sumout + 12 * N;
      }
    }
    if (ensure(ctx, &b->d_items, &b->items_cap, nitem)) return SOSGPU_ERR_CUDA;
    if (ensure(ctx, &b->d_ksets, &b->ksets_cap, nk)) return SOSGPU_ERR_CUDA;
    if (ensure(ctx, &b->d_item_of, &b->item_of_cap, item_of.size())) return SOSGPU_ERR_CUDA;
    if (nitem > b->list_cap) {
      sos_dfree(ctx, b->d_list[0]); sos_dfree(ctx, b->d_list[1]); b->d_list[0] = b->d_list[1] = nullptr; b->list_cap = 0;
      CK(sos_dmalloc(ctx, &b->d_list[0], nitem * sizeof(int)));
      CK(sos_dmalloc(ctx, &b->d_list[1], nitem * sizeof(int)));
      b->list_cap = nitem;
    }
    CK(cudaMemcpyAsync(b->d_items, items.data(), nitem * sizeof(ItemDev), cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(b->d_ksets, ksets.data(), nk * sizeof(KsetDev), cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(b->d_item_of, item_of.data(), item_of.size() * sizeof(int), cudaMemcpyHostToDevice, st));
    {
      std::vector<int> iota(nitem);
      for (size_t i = 0; i < nitem; ++i) iota[i] = (int)i;
      CK(cudaMemcpyAsync(b->d_list[0], iota.data(), nitem * sizeof(int), cudaMemcpyHostToDevice, st));
      CK(cudaStreamSynchronize(st));                             // iota / items are stack-lifetime host buffers
    }

    cudaEvent_t ph[5] = {};
    if (ctx->trace) { for (auto &e : ph) cudaEventCreate(&e); cudaEventRecord(ph[0], st); }
    // ---- kernel matrices of the wave ----
    sos_launch_basis(b->d_ksets, b->d_optics, (int)nk, st);
    sos_launch_kernels(b->d_ksets, b->d_optics, (int)nk, b->maxW, st);
    sos_launch_pack(b->d_ksets, b->d_optics, (int)nk, b->maxKP, st);
    ctx->launches += 3;
    if (ctx->trace) cudaEventRecord(ph[1], st);
    // ---- order 1 ----
    {
      const int nl = sos_launch_order1(b->d_items, b->d_terms, b->d_optics, b->d_ksets, (int)nitem, b->maxKP, (b->maxW - 1) / 2,
                                       b->any_fresnel, st);
      if (nl < 0) { ctx->err = "order-1 kernel launch failed"; return SOSGPU_ERR_CUDA; }
      ctx->launches += nl;
    }
    sos_launch_init(b->d_items, b->d_terms, b->d_optics, (int)nitem, st);
    ctx->launches += 1;
    CK(cudaGetLastError());

    if (ctx->trace) cudaEventRecord(ph[2], st);
    // ---- scattering orders ----
    // The loop never waits for the GPU: the sweep kernel and k_test read the number of still-active items from device
    // memory; the host only follows it with a lag of SOS_LAG orders (pinned counters + events) to shrink k_test's grid
    // and to stop launching once everything has converged.
    int igmax = 0;
    for (int ti : act) igmax = std::max(igmax, b->ho[b->ht[ti].optics].igmax);
    enum { SOS_LAG = 3, SOS_RING = 8 };
    int cur = 0, ub = (int)nitem;
    std::vector<int> order_ev;                                   // event pair index of every launched order
    for (int ig = 2; ig <= igmax; ++ig) {
      if (ig - 2 >= SOS_LAG) {                                   // count after order ig-SOS_LAG
        const int k = (ig - SOS_LAG) % SOS_RING;
        CK(cudaEventSynchronize(ctx->ev_cnt[k]));
        ub = std::min(ub, ctx->h_count[k]);
        if (ub <= 0) break;
      }
      const bool first = (ig == 2);
      const int *list_cur = first ? nullptr : b->d_list[cur];
      const int *count_cur = first ? nullptr : b->d_count + cur;
      CK(cudaMemsetAsync(b->d_count + (cur ^ 1), 0, sizeof(int), st));
      if (b->ev_order.size() < 2 * (order_ev.size() + 1)) {
        cudaEvent_t e0, e1;
        CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
        b->ev_order.push_back(e0); b->ev_order.push_back(e1);
      }
      cudaEvent_t e0 = b->ev_order[2 * order_ev.size()], e1 = b->ev_order[2 * order_ev.size() + 1];
      order_ev.push_back(ig);
      CK(cudaEventRecord(e0, st));
      const int nl = sos_launch_sweep(b->d_items, b->d_terms, b->d_optics, b->d_ksets, list_cur, count_cur, ub, b->maxHB,
                                      ctx->d_work_counter, ctx->num_sms, jdump_dev, st);
      if (nl < 0) { ctx->err = "k_sweep: cannot opt in to its shared-memory size on this device"; return SOSGPU_ERR_CUDA; }
      CK(cudaEventRecord(e1, st));
      sos_launch_test(b->d_items, b->d_terms, b->d_optics, list_cur, count_cur, ub, b->d_list[cur ^ 1], b->d_count + (cur ^ 1), st);
      ctx->launches += nl + 1;
      b->stats.step_launches += nl;
      CK(cudaMemcpyAsync(ctx->h_count + (ig % SOS_RING), b->d_count + (cur ^ 1), sizeof(int), cudaMemcpyDeviceToHost, st));
      CK(cudaEventRecord(ctx->ev_cnt[ig % SOS_RING], st));
      cur ^= 1;
    }
    CK(cudaStreamSynchronize(st));
    CK(cudaGetLastError());
    for (size_t k = 0; k < order_ev.size(); ++k) {
      float ms = 0.f;
      CK(cudaEventElapsedTime(&ms, b->ev_order[2 * k], b->ev_order[2 * k + 1]));
      b->stats.step_ms += ms;
      if (ctx->trace) fprintf(stderr, "wave s0=%d ws=%d ig=%d step_ms=%.3f\n", s0, ws, order_ev[k], ms);
    }

    if (ctx->trace) cudaEventRecord(ph[3], st);
    // ---- per-order bookkeeping and Fourier stop, in order ----
    sos_launch_fourier(b->d_items, b->d_terms, b->d_optics, nterm, b->d_item_of, s0, s1, b->rs_dev, b->w_dev, b->d_rec,
                       b->d_nf, b->d_nsc, b->d_rsn, b->d_emoins, b->d_eplus, b->d_done, st);
    ctx->launches += 1;
    CK(cudaMemcpyAsync(done.data(), b->d_done, nterm * sizeof(int), cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(items.data(), b->d_items, nitem * sizeof(ItemDev), cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    CK(cudaGetLastError());
    if (ctx->trace) {
      cudaEventRecord(ph[4], st); cudaEventSynchronize(ph[4]);
      float m[4];
      for (int k = 0; k < 4; ++k) cudaEventElapsedTime(&m[k], ph[k], ph[k + 1]);
      fprintf(stderr, "wave s0=%d ws=%d items=%zu ksets=%zu: setup kernels %.3f ms, order 1 + init %.3f ms, orders >= 2 %.3f ms, fourier + readback %.3f ms\n",
              s0, ws, nitem, nk, m[0], m[1], m[2], m[3]);
      for (auto &e : ph) cudaEventDestroy(e);
    }
    for (size_t i = 0; i < nitem; ++i) {
      const HostTerm &ht = b->ht[items[i].term];
      const HostOptics &ho = b->ho[ht.optics];
      const long long steps = std::max(0, items[i].n - 1);
      b->stats.steps += steps;
      const double fl = (double)steps * 2.0 * (6.0 * ho.N) * (6.0 * ho.N) * (ht.nt + 1);
      b->stats.flops += fl;
      b->stats.bytes += (double)steps * 96.0 * ho.N * (ht.nt + 1);
      if (steps > 0) item_flops.push_back({items[i].term, items[i].is, fl});
    }
    s0 = s1;
  }

  CK(cudaEventRecord(b->ev0, st));
  sos_launch_aggregate(b->d_terms, b->d_group_start, b->d_group_terms, b->ngroup, b->d_rec, b->d_nf, b->rs_dev, b->w_dev,
                       b->d_grec, b->d_gnrec, st);
  CK(cudaEventRecord(b->ev1, st));
  ctx->launches += 1;
  CK(cudaEventRecord(b->evt1, st));
  CK(cudaStreamSynchronize(st));
  CK(cudaGetLastError());
  float ms = 0.f;
  CK(cudaEventElapsedTime(&ms, b->evt0, b->evt1));
  b->stats.total_ms = ms;
  CK(cudaEventElapsedTime(&ms, b->ev0, b->ev1));
  b->stats.aggregate_ms = ms;
  b->stats.launches = ctx->launches;
  {                                                              // work on Fourier orders the terms kept
    std::vector<int> nf(nterm);
    CK(cudaMemcpy(nf.data(), b->d_nf, nterm * sizeof(int), cudaMemcpyDeviceToHost));
    for (const ItemFlops &f : item_flops) if (f.is < nf[f.term]) b->stats.useful_flops += f.flops;
  }
  return SOSGPU_OK;
}

static int download(sosgpu_ctx *ctx, sosgpu_batch *b, int rec_stride, int wmax, int part_only,
                    sosgpu_term_out *to, sosgpu_group_out *go)
{
  const int nterm = b->nterm, ngroup = b->ngroup;
  const size_t per = (size_t)b->rs_dev * 3 * b->w_dev;
  std::vector<double> tmp;
  auto convert = [&](const double *src, double *dst, int n) {
    // device layout [n][rs_dev][3][w_dev] -> caller layout [n][rec_stride][3][wmax]
    for (int i = 0; i < n; ++i)
      for (int s = 0; s < rec_stride; ++s)
        for (int c = 0; c < 3; ++c) {
          double *d = dst + (((size_t)i * rec_stride + s) * 3 + c) * wmax;
          if (s < b->rs_dev) {
            const double *p = src + (((size_t)i * b->rs_dev + s) * 3 + c) * b->w_dev;
            const int w = std::min(wmax, b->w_dev);
            memcpy(d, p, w * sizeof(double));
            for (int x = w; x < wmax; ++x) d[x] = 0.0;
          } else memset(d, 0, wmax * sizeof(double));
        }
  };
  std::vector<int> nf(nterm);
  CK(cudaMemcpy(nf.data(), b->d_nf, nterm * sizeof(int), cudaMemcpyDeviceToHost));
  std::vector<double> em(nterm), ep(nterm);
  CK(cudaMemcpy(em.data(), b->d_emoins, nterm * sizeof(double), cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(ep.data(), b->d_eplus, nterm * sizeof(double), cudaMemcpyDeviceToHost));
  if (to) {
    if (to->rec) {
      if (rec_stride == b->rs_dev && wmax == b->w_dev)           // caller's layout = device layout: one straight copy
        CK(cudaMemcpy(to->rec, b->d_rec, (size_t)nterm * per * sizeof(double), cudaMemcpyDeviceToHost));
      else {
        tmp.resize((size_t)nterm * per);
        CK(cudaMemcpy(tmp.data(), b->d_rec, tmp.size() * sizeof(double), cudaMemcpyDeviceToHost));
        convert(tmp.data(), to->rec, nterm);
      }
    }
    if (to->n_fourier) memcpy(to->n_fourier, nf.data(), nterm * sizeof(int));
    if (to->n_scatter || to->stop_reason) {
      std::vector<int> a((size_t)nterm * b->rs_dev), r((size_t)nterm * b->rs_dev);
      CK(cudaMemcpy(a.data(), b->d_nsc, a.size() * sizeof(int), cudaMemcpyDeviceToHost));
      CK(cudaMemcpy(r.data(), b->d_rsn, r.size() * sizeof(int), cudaMemcpyDeviceToHost));
      for (int i = 0; i < nterm; ++i)
        for (int s = 0; s < rec_stride; ++s) {
          const bool in = s < b->rs_dev;
          if (to->n_scatter) to->n_scatter[(size_t)i * rec_stride + s] = in ? a[(size_t)i * b->rs_dev + s] : 0;
          if (to->stop_reason) to->stop_reason[(size_t)i * rec_stride + s] = in ? r[(size_t)i * b->rs_dev + s] : -1;
        }
    }
    for (int i = 0; i < nterm; ++i) {
      if (to->emoins) to->emoins[i] = em[i];
      if (to->eplus) to->eplus[i] = ep[i];
      if (to->ttot_tronc) to->ttot_tronc[i] = b->ht[i].ttot_tronc;
      if (to->ttot_vrai) to->ttot_vrai[i] = b->ht[i].ttot_vrai;
      if (to->tauout) to->tauout[i] = b->ht[i].tauout;
      if (to->ier) to->ier[i] = b->ht[i].ier;
    }
  }
  if (go) {
    if (go->rec) {
      if (rec_stride == b->rs_dev && wmax == b->w_dev)
        CK(cudaMemcpy(go->rec, b->d_grec, (size_t)ngroup * per * sizeof(double), cudaMemcpyDeviceToHost));
      else {
        tmp.resize((size_t)ngroup * per);
        CK(cudaMemcpy(tmp.data(), b->d_grec, tmp.size() * sizeof(double), cudaMemcpyDeviceToHost));
        convert(tmp.data(), go->rec, ngroup);
      }
    }
    if (go->n_rec) CK(cudaMemcpy(go->n_rec, b->d_gnrec, ngroup * sizeof(int), cudaMemcpyDeviceToHost));
    // scalars: the caller-owned accumulators of SOS_AGGREGATE.F:452-488, in term order
    for (int g = 0; g < ngroup; ++g) {
      double emg = 0.0, epg = 0.0, tt = 0.0, tv = 0.0, to_ = 0.0;
      for (int x = b->group_start[g]; x < b->group_start[g + 1]; ++x) {
        const int i = b->group_terms[x];
        const double aik = b->ht[i].aik;
        emg = emg + aik * em[i];
        epg = epg + aik * ep[i];
        if (!part_only && b->is_direct(g)) {                     // one solve, no SOS_AGGREGATE: the values of SOS itself
          tt = b->ht[i].ttot_tronc; tv = b->ht[i].ttot_vrai; to_ = b->ht[i].tauout;
        } else if (part_only) {
          tt += aik * std::exp(-b->ht[i].ttot_tronc); tv += aik * std::exp(-b->ht[i].ttot_vrai);
          to_ += aik * std::exp(-b->ht[i].tauout);
        } else {
          double tr;
          tr = (tt != 0) ? aik * std::exp(-b->ht[i].ttot_tronc) + std::exp(-tt) : aik * std::exp(-b->ht[i].ttot_tronc);
          tt = -std::log(tr);
          tr = (tv != 0) ? aik * std::exp(-b->ht[i].ttot_vrai) + std::exp(-tv) : aik * std::exp(-b->ht[i].ttot_vrai);
          tv = -std::log(tr);
          tr = (to_ != 0) ? aik * std::exp(-b->ht[i].tauout) + std::exp(-to_) : aik * std::exp(-b->ht[i].tauout);
          to_ = -std::log(tr);
        }
      }
      if (go->emoins) go->emoins[g] = emg;
      if (go->eplus) go->eplus[g] = epg;
      if (go->ttot_tronc) go->ttot_tronc[g] = tt;
      if (go->ttot_vrai) go->ttot_vrai[g] = tv;
      if (go->tauout) go->tauout[g] = to_;
    }
  }
  return SOSGPU_OK;
}

extern "C" int sosgpu_batch_run(sosgpu_ctx *ctx, sosgpu_batch *b, int rec_stride, int wmax, int part_only,
                                sosgpu_term_out *to, sosgpu_group_out *go)
{
  if (!ctx) return SOSGPU_ERR_NO_DEVICE;
  if (!b) return SOSGPU_ERR_ARG;
  int rc = run_impl(ctx, b, nullptr, 0, nullptr);
  if (rc != SOSGPU_OK) return rc;
  if (to || go) return download(ctx, b, rec_stride, wmax, part_only, to, go);
  return SOSGPU_OK;
}

extern "C" int sosgpu_solve_batch(sosgpu_ctx *ctx, const sosgpu_optics *optics, int noptics,
                                  const sosgpu_term *terms, int nterm, int ngroup,
                                  int rec_stride, int wmax, int part_only,
                                  sosgpu_term_out *term_out, sosgpu_group_out *group_out)
{
  if (!ctx) return SOSGPU_ERR_NO_DEVICE;
  sosgpu_batch *b = nullptr;
  int rc = upload_impl(ctx, optics, noptics, terms, nterm, ngroup, false, &b);
  if (rc != SOSGPU_OK) return rc;
  rc = sosgpu_batch_run(ctx, b, rec_stride, wmax, part_only, term_out, group_out);
  sosgpu_batch_free(ctx, b);
  return rc;
}

extern "C" int sosgpu_group_finalize(double *ttot_tronc, double *ttot_vrai, double *tauout, int ngroup)
{
  for (int g = 0; g < ngroup; ++g) {
    if (ttot_tronc) ttot_tronc[g] = -std::log(ttot_tronc[g]);
    if (ttot_vrai) ttot_vrai[g] = -std::log(ttot_vrai[g]);
    if (tauout) tauout[g] = -std::log(tauout[g]);
  }
  return SOSGPU_OK;
}

extern "C" int sosgpu_batch_set_group_direct(sosgpu_batch *b, const int *direct)
{
  if (!b || !direct) return SOSGPU_ERR_ARG;
  b->group_direct.assign(direct, direct + b->ngroup);
  return SOSGPU_OK;
}

extern "C" int sosgpu_batch_group_buffer(sosgpu_batch *b, void **dev_ptr, size_t *n_doubles)
{
  if (!b || !dev_ptr || !n_doubles) return SOSGPU_ERR_ARG;
  *dev_ptr = b->d_grec;
  *n_doubles = (size_t)b->ngroup * b->rs_dev * 3 * b->w_dev;
  return SOSGPU_OK;
}

extern "C" int sosgpu_batch_stats(const sosgpu_batch *b, sosgpu_stats *st)
{
  if (!b || !st) return SOSGPU_ERR_ARG;
  *st = b->stats;
  return SOSGPU_OK;
}

// ---------------------------------------------------------------------------------------------
// SOS_TRPHI_OPTION over every wavelength of the resident batch (band-solve = term-solves + CKD sum + synthesis).
// Reads the CKD-summed Fourier coefficients where k_aggregate left them (no host round trip).
#include "post_kernels.h"
extern "C" int sosgpu_batch_trphi(sosgpu_ctx *ctx, sosgpu_batch *b, int igli, double wind, double ind_surf, int ifresnel,
                                  int itrphi, double phios, int pas_phi, int ipolar, int nphi_cap, double *up, double *down)
{
  if (!ctx) return SOSGPU_ERR_NO_DEVICE;
  if (!b) return SOSGPU_ERR_ARG;
  CK(cudaSetDevice(ctx->device));
  const double pi = std::acos(-1.0);
  std::vector<double> phis;
  if (itrphi == 1) { phis.push_back(pi + phios * pi / 180.0); phis.push_back(phios * pi / 180.0); }
  else if (itrphi == 2 && pas_phi >= 1) { for (int iphi = 0; iphi <= 360; iphi += pas_phi) phis.push_back(pi * iphi / 180.0); }
  else return SOSGPU_ERR_ARG;
  const int nphi = (int)phis.size(), ng = b->ngroup;
  if (nphi > nphi_cap) return SOSGPU_ERR_ARG;
  std::vector<int> gn(ng);
  CK(cudaMemcpy(gn.data(), b->d_gnrec, ng * sizeof(int), cudaMemcpyDeviceToHost));
  const size_t per = (size_t)b->rs_dev * 3 * b->w_dev;
  int nmax = 0;
  std::vector<TrphiGroup> grp(ng);
  for (int g = 0; g < ng; ++g) {
    double tt = 0.0, to_ = 0.0;
    const int opt = b->group_optics[g];
    if (b->reduced) { tt = b->g_tt[g]; to_ = b->g_to[g]; }       // band-wide values after sosgpu_batch_reduce_groups
    else if (b->is_direct(g)) { const HostTerm &ht = b->ht[b->group_terms[b->group_start[g]]]; tt = ht.ttot_tronc; to_ = ht.tauout; }
    else
      for (int x = b->group_start[g]; x < b->group_start[g + 1]; ++x) {   // SOS_AGGREGATE.F:467-488, term order
        const HostTerm &ht = b->ht[b->group_terms[x]];
        double tr = (tt != 0) ? ht.aik * std::exp(-ht.ttot_tronc) + std::exp(-tt) : ht.aik * std::exp(-ht.ttot_tronc);
        tt = -std::log(tr);
        tr = (to_ != 0) ? ht.aik * std::exp(-ht.tauout) + std::exp(-to_) : ht.aik * std::exp(-ht.tauout);
        to_ = -std::log(tr);
      }
    if (opt < 0) { grp[g] = TrphiGroup{b->d_grec, b->optics_dev[0].rmu, 0, b->ho[0].N, 1, b->w_dev, 0.0, 0.0}; nmax = std::max(nmax, b->ho[0].N); continue; }
    const HostOptics &ho = b->ho[opt];
    if (ho.n0 < 1) { ctx->err = "sosgpu_batch_trphi needs the solar angle among the Gauss angles (n0 > 0)"; return SOSGPU_ERR_ARG; }
    grp[g] = TrphiGroup{b->d_grec + (size_t)g * per, b->optics_dev[opt].rmu, gn[g], ho.N, ho.n0, b->w_dev, tt, to_};
    nmax = std::max(nmax, ho.N);
  }
  for (int i = 0; i < b->noptics; ++i) nmax = std::max(nmax, b->ho[i].N);   // same table pitch on every rank
  // note: rmu[N] on the device holds mu_s (index 0), which SOS_TRPHI never reads
  const size_t nout = (size_t)ng * 2 * 7 * nphi * nmax;
  if (!b->d_tg) CK(sos_dmalloc(ctx, &b->d_tg, ng * sizeof(TrphiGroup)));
  if ((size_t)nphi > b->tphi_cap) { sos_dfree(ctx, b->d_tphi); b->d_tphi = nullptr; CK(sos_dmalloc(ctx, &b->d_tphi, nphi * 8)); b->tphi_cap = nphi; }
  if (nout > b->tout_cap) { sos_dfree(ctx, b->d_tout); b->d_tout = nullptr; CK(sos_dmalloc(ctx, &b->d_tout, nout * 8)); b->tout_cap = nout; }
  TrphiGroup *d_g = (TrphiGroup *)b->d_tg; double *d_phi = b->d_tphi, *d_out = b->d_tout;
  CK(cudaMemcpyAsync(d_g, grp.data(), ng * sizeof(TrphiGroup), cudaMemcpyHostToDevice, ctx->stream));
  CK(cudaMemcpyAsync(d_phi, phis.data(), nphi * 8, cudaMemcpyHostToDevice, ctx->stream));
  CK(cudaMemsetAsync(d_out, 0, nout * 8, ctx->stream));
  const TrphiParams prm = sos_trphi_params(ctx, igli, ifresnel, ipolar, wind, ind_surf, pi);
  cudaEventRecord(ctx->ev_a, ctx->stream);
  sos_launch_trphi_stride(d_g, ng, d_phi, nphi, nmax, prm, d_out, ctx->stream);
  cudaEventRecord(ctx->ev_b, ctx->stream);
  ctx->launches += 1;
  if (up || down) {
    std::vector<double> &out = b->h_tout;
    out.resize(nout);
    CK(cudaMemcpyAsync(out.data(), d_out, nout * 8, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    for (int g = 0; g < ng; ++g)
      for (int ud = 0; ud < 2; ++ud) {
        double *dst = ud == 0 ? up : down;
        if (!dst) continue;
        for (int t = 0; t < 7; ++t)
          for (int ip = 0; ip < nphi; ++ip)
            memcpy(dst + (((size_t)g * 7 + t) * nphi_cap + ip) * nmax, &out[((((size_t)g * 2 + ud) * 7 + t) * nphi + ip) * nmax], nmax * 8);
      }
  } else CK(cudaStreamSynchronize(ctx->stream));
  CK(cudaGetLastError());
  cudaEventElapsedTime(&ctx->last_kernel_ms, ctx->ev_a, ctx->ev_b);
  return nphi;
}

// ---------------------------------------------------------------------------------------------
// single-routine operators
extern "C" int sosgpu_noyaux(sosgpu_ctx *ctx, int is, int nbmu, const double *rmu, int os_nb,
                             const double *alpha, const double *beta, const double *gamma, const double *zeta,
                             double *xpl, double *xrl, double *xtl,
                             double *bp, double *gr, double *gt, double *arr, double *art, double *att)
{
  if (!ctx) return SOSGPU_ERR_NO_DEVICE;
  if (nbmu < 1 || nbmu > SOSGPU_NBMU_MAX || os_nb < 2 || os_nb > SOSGPU_NB_MAX || is < 0) return SOSGPU_ERR_ARG;
  CK(cudaSetDevice(ctx->device));
  const int N = nbmu, W = 2 * N + 1;
  Arena ar;
  size_t o_rmu = ar.put(rmu, W * 8), o_a = ar.put(alpha, (os_nb + 1) * 8), o_b = ar.put(beta, (os_nb + 1) * 8);
  size_t o_g = ar.put(gamma, (os_nb + 1) * 8), o_z = ar.put(zeta, (os_nb + 1) * 8);
  const size_t nb_basis = (size_t)3 * (os_nb + 2) * W, nb_ker = (size_t)6 * W * W, nb_xpl = (size_t)3 * W;
  char *d = nullptr; double *dw = nullptr; OpticsDev *dop = nullptr; KsetDev *dks = nullptr;
  SosFreeGuard guard(ctx);                                       // releases the temporaries on every return path
  CK(sos_dmalloc(ctx, &d, ar.size())); guard.add(d);
  CK(sos_dmalloc(ctx, &dw, (nb_basis + nb_ker + nb_xpl) * 8)); guard.add(dw);
  CK(sos_dmalloc(ctx, &dop, sizeof(OpticsDev))); guard.add(dop);
  CK(sos_dmalloc(ctx, &dks, sizeof(KsetDev))); guard.add(dks);
  CK(cudaMemcpy(d, ar.data(), ar.size(), cudaMemcpyHostToDevice));
  CK(cudaMemset(dw, 0, (nb_basis + nb_ker + nb_xpl) * 8));
  OpticsDev op{};
  op.nbmu = N; op.W = W; op.HB = roundup(3 * N, 16); op.KP = 2 * op.HB; op.os_nb = os_nb;
  op.rmu = (const double *)(d + o_rmu); op.alpha = (const double *)(d + o_a); op.beta = (const double *)(d + o_b);
  op.gamma = (const double *)(d + o_g); op.zeta = (const double *)(d + o_z);
  KsetDev ks{};
  ks.optics = 0; ks.is = is; ks.basis = dw; ks.ker = dw + nb_basis; ks.xpl = dw + nb_basis + nb_ker;
  CK(cudaMemcpy(dop, &op, sizeof(op), cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dks, &ks, sizeof(ks), cudaMemcpyHostToDevice));
  sos_launch_basis(dks, dop, 1, ctx->stream);
  sos_launch_kernels(dks, dop, 1, W, ctx->stream);
  ctx->launches += 2;
  CK(cudaStreamSynchronize(ctx->stream));
  CK(cudaGetLastError());
  std::vector<double> ker(nb_ker), x3(nb_xpl);
  CK(cudaMemcpy(ker.data(), ks.ker, nb_ker * 8, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(x3.data(), ks.xpl, nb_xpl * 8, cudaMemcpyDeviceToHost));
  const size_t WW = (size_t)W * W;
  memcpy(bp, &ker[0], WW * 8); memcpy(gr, &ker[WW], WW * 8); memcpy(gt, &ker[2 * WW], WW * 8);
  memcpy(arr, &ker[3 * WW], WW * 8); memcpy(art, &ker[4 * WW], WW * 8); memcpy(att, &ker[5 * WW], WW * 8);
  memcpy(xpl, &x3[0], W * 8); memcpy(xrl, &x3[W], W * 8); memcpy(xtl, &x3[2 * W], W * 8);
  return SOSGPU_OK;
}

extern "C" int sosgpu_order_step(sosgpu_ctx *ctx, int is, int nbmu, const double *rmu, const double *ga, int os_nb,
                                 const double *alpha, const double *beta, const double *gamma, const double *zeta,
                                 double ron, int ipolar, int nt, const double *h, const double *xdel, const double *ydel,
                                 const double *i1, const double *q1, const double *u1, const double *bc_unused,
                                 double *i1n, double *q1n, double *u1n, double *i2, double *q2, double *u2)
{
  (void)bc_unused;
  if (!ctx) return SOSGPU_ERR_NO_DEVICE;
  CK(cudaSetDevice(ctx->device));
  const int N = nbmu, W = 2 * N + 1, L = nt + 1;
  // one optics, one term at SOS_OS level (no truncation adaptation), one item
  sosgpu_optics o{};
  o.nbmu = N; o.rmu = rmu; o.ga = ga; o.n0 = 0; o.tetas = 0.0; o.os_nb = os_nb; o.alpha = alpha; o.beta = beta;
  o.gamma = gamma; o.zeta = zeta; o.ron = ron; o.rho = 0.0; o.imat_surf = 0; o.ifresnel = 0; o.ind_surf = 1.34;
  o.igmax = 100; o.ipolar = ipolar; o.zout = -1.0; o.piz = 1.0; o.piztr = 1.0;
  std::vector<double> rmu2(rmu, rmu + W);
  // keep the caller's mu_s (index 0) as the solar direction: n0<=0 -> tetas from rmu(0) = -cos(tetas)
  o.tetas = std::acos(-rmu2[N]) * 180.0 / std::acos(-1.0);
  std::vector<double> zprof(L, 0.0);
  sosgpu_term t{};
  t.optics = 0; t.group = 0; t.aik = 1.0; t.nt = nt; t.zprof = zprof.data(); t.h = h; t.pcaer = xdel; t.pcmol = ydel;
  sosgpu_batch *b = nullptr;
  int rc = upload_impl(ctx, &o, 1, &t, 1, 1, true, &b);
  if (rc != SOSGPU_OK) return rc;
  struct BatchGuard { sosgpu_ctx *c; sosgpu_batch *b; ~BatchGuard() { sosgpu_batch_free(c, b); } } bguard{ctx, b};
  SosFreeGuard guard(ctx);
  b->optics_dev[0].tab = rmu2[N];       // exact mu_s of the caller
  // rmu[N] in the arena was set from tetas; overwrite with the exact value
  CK(cudaMemcpy((void *)(b->optics_dev[0].rmu + N), &rmu2[N], 8, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(b->d_optics, b->optics_dev.data(), sizeof(OpticsDev), cudaMemcpyHostToDevice));
  const HostOptics &ho = b->ho[0];
  const HostTerm &ht = b->ht[0];
  const int KP = ho.KP, HB = ho.HB, LP = ht.LP;
  // kernel set
  const size_t al = 256;
  auto rup = [&](size_t x) { return (x + al - 1) / al * al; };
  KsetDev ks{};
  ks.optics = 0; ks.is = is; ks.dual = (is <= 2) ? 1 : 0; ks.beta0 = (is == 0) ? 1.0 : 0.0;
  const size_t kbytes = rup(3 * (os_nb + 2) * W * 8) + rup(6 * W * W * 8) + rup(3 * W * 8) + 2 * rup((size_t)KP * KP * 8) + 5 * rup(KP * 8) + rup(16 * KP * 8);
  char *kp = nullptr;
  CK(sos_dmalloc(ctx, &kp, kbytes)); guard.add(kp);
  CK(cudaMemset(kp, 0, kbytes));
  char *p = kp;
  ks.basis = (double *)p; p += rup(3 * (os_nb + 2) * W * 8);
  ks.ker = (double *)p; p += rup(6 * W * W * 8);
  ks.xpl = (double *)p; p += rup(3 * W * 8);
  ks.apackA = (double *)p; p += rup((size_t)KP * KP * 8);
  ks.apackR = (double *)p; p += rup((size_t)KP * KP * 8);
  ks.c1 = (double *)p; p += rup(KP * 8);
  ks.c2 = (double *)p; p += rup(KP * 8);
  ks.fz1 = (double *)p; p += rup(KP * 8);
  ks.fz2 = (double *)p; p += rup(KP * 8);
  ks.urow = (double *)p; p += rup(KP * 8);
  ks.vpack = (double *)p; p += rup(16 * KP * 8);
  KsetDev *dks = nullptr;
  CK(sos_dmalloc(ctx, &dks, sizeof(KsetDev))); guard.add(dks);
  CK(cudaMemcpy(dks, &ks, sizeof(ks), cudaMemcpyHostToDevice));
  // fields: x[1] = input (order n=1 parity), x[0] = output, plus J dump
  const size_t fsz = std::max(SOS_XSIZE(KP, L), (size_t)KP * LP);   // field (chunk-major) or J dump ([row][LP])
  std::vector<double> xin(fsz, 0.0);
  const double *src[3] = {i1, q1, u1};
  for (int d = 0; d < 2; ++d)
    for (int s = 0; s < 3; ++s)
      for (int k = 1; k <= N; ++k) {
        const int r = d * HB + s * N + (k - 1);
        const int kk = d == 0 ? k : -k;
        for (int lv = 0; lv < L; ++lv) xin[SOS_XIDX(KP, r, lv)] = src[s][(size_t)(kk + N) * L + lv];
      }
  double *dx = nullptr;
  CK(sos_dmalloc(ctx, &dx, 3 * fsz * 8)); guard.add(dx);
  CK(cudaMemset(dx, 0, 3 * fsz * 8));
  CK(cudaMemcpy(dx + fsz, xin.data(), fsz * 8, cudaMemcpyHostToDevice));
  ItemDev it{};
  it.term = 0; it.is = is; it.kset = 0; it.n = 1; it.active = 1;
  it.x[0] = dx; it.x[1] = dx + fsz;
  ItemDev *dit = nullptr;
  CK(sos_dmalloc(ctx, &dit, sizeof(ItemDev))); guard.add(dit);
  CK(cudaMemcpy(dit, &it, sizeof(it), cudaMemcpyHostToDevice));
  sos_launch_basis(dks, b->d_optics, 1, ctx->stream);
  sos_launch_kernels(dks, b->d_optics, 1, W, ctx->stream);
  sos_launch_pack(dks, b->d_optics, 1, KP, ctx->stream);
  ctx->launches += 3;
  ctx->launches += sos_launch_sweep(dit, b->d_terms, b->d_optics, dks, nullptr, nullptr, 1, HB, ctx->d_work_counter, ctx->num_sms,
                                    dx + 2 * fsz, ctx->stream);
  CK(cudaStreamSynchronize(ctx->stream));
  CK(cudaGetLastError());
  std::vector<double> xo(fsz), jo(fsz);
  CK(cudaMemcpy(xo.data(), dx, fsz * 8, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(jo.data(), dx + 2 * fsz, fsz * 8, cudaMemcpyDeviceToHost));
  double *dstx[3] = {i1n, q1n, u1n}, *dstj[3] = {i2, q2, u2};
  for (int d = 0; d < 2; ++d)
    for (int s = 0; s < 3; ++s)
      for (int k = 1; k <= N; ++k) {
        const int r = d * HB + s * N + (k - 1);
        const int kk = d == 0 ? k : -k;
        for (int lv = 0; lv < L; ++lv) {
          if (dstx[s]) dstx[s][(size_t)(kk + N) * L + lv] = xo[SOS_XIDX(KP, r, lv)];
          if (dstj[s]) dstj[s][(size_t)(kk + N) * L + lv] = jo[(size_t)r * LP + lv];
        }
      }
  return SOSGPU_OK;
}
