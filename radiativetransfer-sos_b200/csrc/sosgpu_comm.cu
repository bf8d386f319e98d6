// sosgpu_comm.cu -- the multi-GPU part of libsosgpu.so: one NCCL communicator per context (one process per GPU), owned by the
// library (no PyTorch, no MPI needed for the data path; the 128-byte unique id travels by whatever the host program has:
// MPI_Bcast, a file, torch.distributed).  NCCL is bound at run time (dlopen of libnccl.so.2), so a process that already
// carries an NCCL (e.g. PyTorch's) shares it instead of loading a second one.
//
// Two layouts:
//   * term-sharded   every rank solves a subset of the CKD terms of every wavelength; the partial CKD-weighted sums
//                    [ngroup][S+1][3][W] + per-group scalars + series-length indicators are ONE contiguous device buffer,
//                    reduced in place by ONE ncclReduce (sum, f64) to the root (SOS_AGGREGATE.F:351-488 across GPUs);
//   * wavelength-sharded  every rank owns whole wavelengths (no reduce); sosgpu_batch_gather_tables collects the
//                    synthesised SOS_TRPHI tables of all ranks on the root with one grouped ncclSend/ncclRecv.
#include "sosgpu_host.h"
#include "post_kernels.h"

#include <dlfcn.h>
#include <nccl.h>

#include <cmath>
#include <cstring>
#include <mutex>

namespace {

struct NcclApi {
  void *handle = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*Reduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*AllReduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Send)(const void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Recv)(void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  const char *(*GetErrorString)(ncclResult_t) = nullptr;
};

NcclApi *nccl_api(std::string &err)
{
  static NcclApi api;
  static std::once_flag once;
  static std::string load_err;
  std::call_once(once, [] {
    const char *names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char *n : names) {
      api.handle = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
      if (api.handle) break;
    }
    if (!api.handle) { load_err = std::string("cannot load NCCL: ") + dlerror(); return; }
#define SYM(field, name)                                                                 \
  *(void **)(&api.field) = dlsym(api.handle, name);                                      \
  if (!api.field) { load_err = std::string("NCCL symbol missing: ") + name; return; }
    SYM(GetUniqueId, "ncclGetUniqueId") SYM(CommInitRank, "ncclCommInitRank") SYM(CommDestroy, "ncclCommDestroy")
    SYM(Reduce, "ncclReduce") SYM(AllReduce, "ncclAllReduce") SYM(Send, "ncclSend") SYM(Recv, "ncclRecv")
    SYM(GroupStart, "ncclGroupStart") SYM(GroupEnd, "ncclGroupEnd") SYM(GetErrorString, "ncclGetErrorString")
#undef SYM
  });
  if (!load_err.empty()) { err = load_err; return nullptr; }
  return &api;
}

#define NK(call)                                                                                   \
  do {                                                                                             \
    ncclResult_t r_ = (call);                                                                      \
    if (r_ != ncclSuccess) {                                                                       \
      ctx->err = std::string(#call) + ": " + api->GetErrorString(r_);                              \
      return SOSGPU_ERR_CUDA;                                                                      \
    }                                                                                              \
  } while (0)

}  // namespace

static_assert(SOSGPU_UNIQUE_ID_BYTES == sizeof(ncclUniqueId), "ncclUniqueId size");

extern "C" int sosgpu_comm_unique_id(char *id)
{
  if (!id) return SOSGPU_ERR_ARG;
  std::string err;
  NcclApi *api = nccl_api(err);
  if (!api) return SOSGPU_ERR_CUDA;
  ncclUniqueId u;
  if (api->GetUniqueId(&u) != ncclSuccess) return SOSGPU_ERR_CUDA;
  memcpy(id, &u, sizeof(u));
  return SOSGPU_OK;
}

extern "C" int sosgpu_comm_init(sosgpu_ctx *ctx, int nranks, int rank, const char *id)
{
  if (!ctx) return SOSGPU_ERR_NO_DEVICE;
  if (!id || nranks < 1 || rank < 0 || rank >= nranks) { ctx->err = "bad communicator arguments"; return SOSGPU_ERR_ARG; }
  NcclApi *api = nccl_api(ctx->err);
  if (!api) return SOSGPU_ERR_CUDA;
  CK(cudaSetDevice(ctx->device));
  if (ctx->nccl_comm) { api->CommDestroy((ncclComm_t)ctx->nccl_comm); ctx->nccl_comm = nullptr; }
  ncclUniqueId u;
  memcpy(&u, id, sizeof(u));
  ncclComm_t comm = nullptr;
  NK(api->CommInitRank(&comm, nranks, u, rank));
  ctx->nccl_comm = comm; ctx->nranks = nranks; ctx->rank = rank;
  return SOSGPU_OK;
}

extern "C" int sosgpu_comm_destroy(sosgpu_ctx *ctx)
{
  if (!ctx) return SOSGPU_ERR_NO_DEVICE;
  if (!ctx->nccl_comm) return SOSGPU_OK;
  NcclApi *api = nccl_api(ctx->err);
  if (!api) return SOSGPU_ERR_CUDA;
  cudaSetDevice(ctx->device);
  cudaStreamSynchronize(ctx->stream);
  api->CommDestroy((ncclComm_t)ctx->nccl_comm);
  ctx->nccl_comm = nullptr; ctx->nranks = 1; ctx->rank = 0;
  return SOSGPU_OK;
}

extern "C" int sosgpu_comm_barrier(sosgpu_ctx *ctx)
{
  if (!ctx) return SOSGPU_ERR_NO_DEVICE;
  if (!ctx->nccl_comm) return SOSGPU_OK;
  NcclApi *api = nccl_api(ctx->err);
  if (!api) return SOSGPU_ERR_CUDA;
  CK(cudaSetDevice(ctx->device));
  NK(api->AllReduce(ctx->d_work_counter + 8, ctx->d_work_counter + 8, 1, ncclInt32, ncclSum, (ncclComm_t)ctx->nccl_comm, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  return SOSGPU_OK;
}

// Tail of the group-sum buffer, per group: 8 scalars {sum a*EMOINS, sum a*EPLUS, sum a*exp(-TTOT_TRONC), sum a*exp(-TTOT_VRAI),
// sum a*exp(-TAUOUT), 0, 0, 0} then rs_dev indicators (1.0 for every Fourier order this rank holds a record of): the
// reduced indicators count the ranks, the longest series of the band is the number of non-zero entries.
extern "C" int sosgpu_batch_reduce_groups(sosgpu_ctx *ctx, sosgpu_batch *b, int root)
{
  if (!ctx) return SOSGPU_ERR_NO_DEVICE;
  if (!b || root < 0 || root >= ctx->nranks) { ctx->err = "bad reduce arguments"; return SOSGPU_ERR_ARG; }
  CK(cudaSetDevice(ctx->device));
  const int ng = b->ngroup, nterm = b->nterm, rs = b->rs_dev;
  const size_t per = (size_t)rs * 3 * b->w_dev, tail_per = 8 + (size_t)rs;
  // local partial scalars, in term order (SOS_AGGREGATE.F:452-488 without the running -log)
  std::vector<double> em(nterm), ep(nterm);
  std::vector<int> gn(ng);
  CK(cudaMemcpyAsync(em.data(), b->d_emoins, nterm * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaMemcpyAsync(ep.data(), b->d_eplus, nterm * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaMemcpyAsync(gn.data(), b->d_gnrec, ng * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  std::vector<double> tail((size_t)ng * tail_per, 0.0);
  for (int g = 0; g < ng; ++g) {
    double *t = &tail[(size_t)g * tail_per];
    for (int x = b->group_start[g]; x < b->group_start[g + 1]; ++x) {
      const int i = b->group_terms[x];
      const HostTerm &ht = b->ht[i];
      if (ht.ier != 0) continue;
      t[0] = t[0] + ht.aik * em[i];
      t[1] = t[1] + ht.aik * ep[i];
      t[2] += ht.aik * std::exp(-ht.ttot_tronc);
      t[3] += ht.aik * std::exp(-ht.ttot_vrai);
      t[4] += ht.aik * std::exp(-ht.tauout);
    }
    for (int s = 0; s < gn[g] && s < rs; ++s) t[8 + s] = 1.0;
  }
  double *d_tail = b->d_grec + (size_t)ng * per;
  CK(cudaMemcpyAsync(d_tail, tail.data(), tail.size() * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
  const size_t count = (size_t)ng * (per + tail_per);
  if (ctx->nccl_comm && ctx->nranks > 1) {
    NcclApi *api = nccl_api(ctx->err);
    if (!api) return SOSGPU_ERR_CUDA;
    NK(api->Reduce(b->d_grec, b->d_grec, count, ncclDouble, ncclSum, root, (ncclComm_t)ctx->nccl_comm, ctx->stream));
  }
  if (ctx->rank == root) {
    CK(cudaMemcpyAsync(tail.data(), d_tail, tail.size() * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    b->g_em.resize(ng); b->g_ep.resize(ng); b->g_tt.resize(ng); b->g_tv.resize(ng); b->g_to.resize(ng); b->g_nrec.resize(ng);
    for (int g = 0; g < ng; ++g) {
      const double *t = &tail[(size_t)g * tail_per];
      b->g_em[g] = t[0]; b->g_ep[g] = t[1];
      b->g_tt[g] = t[2] > 0 ? -std::log(t[2]) : 0.0;             // SOS_AGGREGATE.F:467-488 (closed form of the running -log)
      b->g_tv[g] = t[3] > 0 ? -std::log(t[3]) : 0.0;
      b->g_to[g] = t[4] > 0 ? -std::log(t[4]) : 0.0;
      int n = 0;
      for (int s = 0; s < rs; ++s) if (t[8 + s] != 0.0) n = s + 1;
      b->g_nrec[g] = n;
    }
    CK(cudaMemcpyAsync(b->d_gnrec, b->g_nrec.data(), ng * sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
    b->reduced = true;
  }
  CK(cudaStreamSynchronize(ctx->stream));
  return SOSGPU_OK;
}

extern "C" int sosgpu_batch_set_group_optics(sosgpu_batch *b, const int *optics_of_group)
{
  if (!b || !optics_of_group) return SOSGPU_ERR_ARG;
  for (int g = 0; g < b->ngroup; ++g)
    if (optics_of_group[g] < 0 || optics_of_group[g] >= b->noptics) return SOSGPU_ERR_ARG;
  b->group_optics.assign(optics_of_group, optics_of_group + b->ngroup);
  return SOSGPU_OK;
}

// Band sums after sosgpu_batch_reduce_groups (root only): what the SOS_AGGREGATE chain leaves behind for every wavelength.
extern "C" int sosgpu_batch_groups(sosgpu_ctx *ctx, sosgpu_batch *b, int rec_stride, int wmax, sosgpu_group_out *go)
{
  if (!ctx) return SOSGPU_ERR_NO_DEVICE;
  if (!b || !go) return SOSGPU_ERR_ARG;
  if (!b->reduced) { ctx->err = "sosgpu_batch_groups: call sosgpu_batch_reduce_groups first (on the root rank)"; return SOSGPU_ERR_ARG; }
  CK(cudaSetDevice(ctx->device));
  const int ng = b->ngroup;
  const size_t per = (size_t)b->rs_dev * 3 * b->w_dev;
  if (go->rec) {
    std::vector<double> tmp((size_t)ng * per);
    CK(cudaMemcpy(tmp.data(), b->d_grec, tmp.size() * sizeof(double), cudaMemcpyDeviceToHost));
    const int w = std::min(wmax, b->w_dev);
    for (int g = 0; g < ng; ++g)
      for (int s = 0; s < rec_stride; ++s)
        for (int c = 0; c < 3; ++c) {
          double *d = go->rec + (((size_t)g * rec_stride + s) * 3 + c) * wmax;
          memset(d, 0, wmax * sizeof(double));
          if (s < b->rs_dev) memcpy(d, &tmp[(((size_t)g * b->rs_dev + s) * 3 + c) * b->w_dev], w * sizeof(double));
        }
  }
  for (int g = 0; g < ng; ++g) {
    if (go->n_rec) go->n_rec[g] = b->g_nrec[g];
    if (go->emoins) go->emoins[g] = b->g_em[g];
    if (go->eplus) go->eplus[g] = b->g_ep[g];
    if (go->ttot_tronc) go->ttot_tronc[g] = b->g_tt[g];
    if (go->ttot_vrai) go->ttot_vrai[g] = b->g_tv[g];
    if (go->tauout) go->tauout[g] = b->g_to[g];
  }
  return SOSGPU_OK;
}

// Wavelength-sharded layout: the tables the last sosgpu_batch_trphi call left on the device (this rank's groups) are
// collected on the root in rank order.  groups_of_rank[nranks]; up/down (root only, may be NULL for a device-resident
// timing): [sum groups][7][nphi][nmax] with nphi, nmax as returned/used by the synthesis call (identical on every rank).
extern "C" int sosgpu_batch_gather_tables(sosgpu_ctx *ctx, sosgpu_batch *b, int root, const int *groups_of_rank, int nphi, int nmax,
                                          double *up, double *down)
{
  if (!ctx) return SOSGPU_ERR_NO_DEVICE;
  if (!b || !groups_of_rank || root < 0 || root >= ctx->nranks || nphi < 1 || nmax < 1) { ctx->err = "bad gather arguments"; return SOSGPU_ERR_ARG; }
  CK(cudaSetDevice(ctx->device));
  const size_t per_group = (size_t)2 * 7 * nphi * nmax;
  if (groups_of_rank[ctx->rank] != b->ngroup || (size_t)b->ngroup * per_group > b->tout_cap) {
    ctx->err = "sosgpu_batch_gather_tables: run sosgpu_batch_trphi on this batch first"; return SOSGPU_ERR_ARG;
  }
  size_t total = 0, my_off = 0;
  for (int r = 0; r < ctx->nranks; ++r) { if (r == ctx->rank) my_off = total; total += groups_of_rank[r]; }
  const bool is_root = ctx->rank == root;
  double *d_all = b->d_tout;
  if (ctx->nranks > 1) {
    NcclApi *api = nccl_api(ctx->err);
    if (!api || !ctx->nccl_comm) { ctx->err = "no communicator (sosgpu_comm_init)"; return SOSGPU_ERR_ARG; }
    if (is_root) {
      if (total * per_group > ctx->gather_cap) {
        if (ctx->d_gather) cudaFree(ctx->d_gather);
        ctx->d_gather = nullptr; ctx->gather_cap = 0;
        CK(cudaMalloc(&ctx->d_gather, total * per_group * sizeof(double)));
        ctx->gather_cap = total * per_group;
      }
      d_all = ctx->d_gather;
      CK(cudaMemcpyAsync(d_all + my_off * per_group, b->d_tout, (size_t)b->ngroup * per_group * sizeof(double), cudaMemcpyDeviceToDevice, ctx->stream));
    }
    NK(api->GroupStart());
    if (is_root) {
      size_t off = 0;
      for (int r = 0; r < ctx->nranks; ++r) {
        if (r != root && groups_of_rank[r] > 0)
          NK(api->Recv(d_all + off * per_group, (size_t)groups_of_rank[r] * per_group, ncclDouble, r, (ncclComm_t)ctx->nccl_comm, ctx->stream));
        off += groups_of_rank[r];
      }
    } else if (b->ngroup > 0) {
      NK(api->Send(b->d_tout, (size_t)b->ngroup * per_group, ncclDouble, root, (ncclComm_t)ctx->nccl_comm, ctx->stream));
    }
    NK(api->GroupEnd());
  }
  if (is_root && (up || down)) {
    // device [g][2 (up, down)][7][nphi][nmax] -> the two caller tables, one strided copy each (no staging buffer)
    const size_t half = (size_t)7 * nphi * nmax;
    if (up) CK(cudaMemcpy2DAsync(up, half * 8, d_all, per_group * 8, half * 8, total, cudaMemcpyDeviceToHost, ctx->stream));
    if (down) CK(cudaMemcpy2DAsync(down, half * 8, d_all + half, per_group * 8, half * 8, total, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
  } else CK(cudaStreamSynchronize(ctx->stream));
  return SOSGPU_OK;
}
