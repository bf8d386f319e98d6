// sosgpu_host.h -- host-side definitions shared by the translation units of libsosgpu.so
#pragma once
#include "../../include/sosgpu.h"
#include "sosgpu_internal.h"
#include "post_kernels.h"
#include <string>
#include <vector>

#define CK(call)                                                                                   \
  do {                                                                                             \
    cudaError_t e_ = (call);                                                                       \
    if (e_ != cudaSuccess) {                                                                       \
      ctx->err = std::string(#call) + ": " + cudaGetErrorString(e_);                               \
      return SOSGPU_ERR_CUDA;                                                                      \
    }                                                                                              \
  } while (0)

struct sosgpu_ctx {
  int device = 0;
  cudaStream_t stream = nullptr;
  std::string err;
  long long launches = 0;
  size_t field_budget = (size_t)48 << 30;
  int max_wave_orders = 0;
  cudaEvent_t ev_a = nullptr, ev_b = nullptr; float last_kernel_ms = 0.f;   // device time of the last glitter / synthesis kernel
  int *h_count = nullptr;            // pinned ring of active counts, read back with a lag (the wave loop never waits)
  cudaEvent_t ev_cnt[8] = {};        // one event per ring slot
  sosgpu_direct_models dm = {};      // land-surface direct terms of SOS_TRPHI (sosgpu_set_direct_models)
  bool trace = false;                // SOS_TRACE, read once at create
  char *cache_field = nullptr; size_t cache_field_bytes = 0;   // wave pools parked by the last freed batch
  char *cache_kpool = nullptr; size_t cache_kpool_bytes = 0;
  double *grec_cache = nullptr; size_t grec_cache_bytes = 0;   // group-sum buffer parked by the last freed batch
  char *h_arena = nullptr; size_t h_arena_cap = 0;     // pinned staging buffer of the batch uploads
  unsigned *d_work_counter = nullptr; int num_sms = 0;   // work queue of the persistent sweep kernel
  void *nccl_comm = nullptr; int nranks = 1, rank = 0;   // sosgpu_comm_init
  double *d_gather = nullptr; size_t gather_cap = 0; std::vector<double> h_gather;   // root side of sosgpu_batch_gather_tables
  cudaMemPool_t pool = nullptr;      // stream-ordered pool (release threshold = never) behind sos_dmalloc / sos_dfree
};

// the process-wide context of the gfortran-ABI symbols, created on first use (sosgpu_shims.cu); nullptr without a usable device
sosgpu_ctx *sos_shim_ctx();

// frees a list of pool allocations when a routine leaves early (CK returns) or normally
struct SosFreeGuard {
  sosgpu_ctx *ctx; void *p[8]; int n = 0;
  explicit SosFreeGuard(sosgpu_ctx *c) : ctx(c) {}
  void add(void *q) { if (n < 8) p[n++] = q; }
  ~SosFreeGuard();
};

template <class T> static inline cudaError_t sos_dmalloc(sosgpu_ctx *ctx, T **p, size_t bytes)
{
  return cudaMallocFromPoolAsync((void **)p, bytes ? bytes : 8, ctx->pool, ctx->stream);
}
static inline void sos_dfree(sosgpu_ctx *ctx, void *p)
{
  if (!p) return;
  if (ctx) cudaFreeAsync(p, ctx->stream);
  else cudaFree(p);
}
inline SosFreeGuard::~SosFreeGuard() { for (int i = 0; i < n; ++i) sos_dfree(ctx, p[i]); }

struct HostOptics {
  int N, W, HB, KP, os_nb, n0, imat_surf, ifresnel, ipolar, igmax, n_surf_rec;
  double tab, ro, ron, ind_surf, zout, beta2, gamma2, alpha2, f11sun, f12sun, a_trunc, piz, piztr;
  std::vector<double> rmu, ga, alpha, beta, gamma, zeta, f11, f12, f33;
  const float *surf;
  bool limb;
};

struct HostTerm {
  int optics, group, nt, LP, iborm, jout, ier;
  double zz, aik, eground, ttot_vrai, ttot_tronc, tauout;
  std::vector<double> h, xdel, ydel, dt, inv, ch, cf;
};

struct sosgpu_batch {
  std::vector<HostOptics> ho;
  std::vector<HostTerm> ht;
  int nterm = 0, noptics = 0, ngroup = 0;
  int rs_dev = 0, w_dev = 0, maxHB = 0, maxW = 0, maxKP = 0, maxNB = 0, smax = 0, any_fresnel = 0;
  char *d_arena = nullptr;
  OpticsDev *d_optics = nullptr;
  TermDev *d_terms = nullptr;
  std::vector<OpticsDev> optics_dev;
  std::vector<TermDev> terms_dev;
  double *d_att = nullptr, *d_i4 = nullptr;
  size_t i4_total = 0;
  size_t grec_bytes = 0;
  double *d_rec = nullptr, *d_emoins = nullptr, *d_eplus = nullptr, *d_grec = nullptr;
  int *d_nf = nullptr, *d_nsc = nullptr, *d_rsn = nullptr, *d_done = nullptr, *d_gnrec = nullptr;
  int *d_group_start = nullptr, *d_group_terms = nullptr;
  std::vector<int> group_start, group_terms;
  // wave pools (grown on demand)
  char *d_field = nullptr; size_t field_bytes = 0;
  char *d_kpool = nullptr; size_t kpool_bytes = 0;
  ItemDev *d_items = nullptr; size_t items_cap = 0;
  KsetDev *d_ksets = nullptr; size_t ksets_cap = 0;
  int *d_item_of = nullptr; size_t item_of_cap = 0;
  int *d_list[2] = {nullptr, nullptr}; size_t list_cap = 0;
  int *d_count = nullptr;           // [2]
  int *h_count = nullptr;           // pinned
  cudaEvent_t ev0 = nullptr, ev1 = nullptr, evt0 = nullptr, evt1 = nullptr;
  std::vector<cudaEvent_t> ev_order;  // timing event pairs of the sweep launches of a wave
  sosgpu_stats stats{};
  // persistent buffers of sosgpu_batch_trphi
  void *d_tg = nullptr; double *d_tphi = nullptr, *d_tout = nullptr; size_t tout_cap = 0, tphi_cap = 0;
  std::vector<double> h_tout;
  // multi-GPU (sosgpu_comm.cu): optics entry of every group (default: first local term's), and the band-wide group
  // metadata after sosgpu_batch_reduce_groups (sums of a*exp(-tau), fluxes, longest series)
  std::vector<int> group_optics;
  // groups SOS_PROC solves once and does not pass through SOS_AGGREGATE (sosgpu_batch_set_group_direct)
  std::vector<char> group_direct;
  bool is_direct(int g) const { return (size_t)g < group_direct.size() && group_direct[g] && group_start[g + 1] - group_start[g] == 1; }
  bool reduced = false;
  std::vector<double> g_em, g_ep, g_tt, g_tv, g_to;
  std::vector<int> g_nrec;
};


static inline TrphiParams sos_trphi_params(const sosgpu_ctx *ctx, int igli, int ifresnel, int ipolar, double wind, double ind_surf, double pi)
{
  TrphiParams p{};
  p.igli = igli; p.ifresnel = ifresnel; p.ipolar = ipolar; p.wind = wind; p.ind_surf = ind_surf; p.pi = pi;
  const sosgpu_direct_models &d = ctx->dm;
  p.iroujean = d.iroujean; p.k0 = d.k0; p.k1 = d.k1; p.k2 = d.k2;
  p.irondeaux = d.irondeaux; p.ibreon = d.ibreon; p.inadal = d.inadal; p.alpha_nadal = d.alpha_nadal; p.beta_nadal = d.beta_nadal;
  p.imaignan = d.imaignan; p.coef_c_maignan = d.coef_c_maignan;
  return p;
}
