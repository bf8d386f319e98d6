// sweep_kernel.cu -- the hot kernel of libsosgpu.so (sm_100a): one launch advances every active (term, Fourier order)
// item by one scattering order n >= 2.
//
//   J(level, mu) = XDEL(level) * (A_A X_{n-1})(level) + YDEL(level) * (A_R X_{n-1})(level)
//        = SOS_FSOURCE_ORDREIG (SOS_OS.F:2663-3017) as a dense FP64 contraction on DMMA tiles (mma.sync.m8n8k4.f64),
//   X_n = SOS_INTEGR_EPOPT(J, boundary values) (SOS_OS.F:2222-2357) on the tile while it is on chip; boundary values at
//        the ground (Lambert / BRDF-BPDF quadrature / flat Fresnel, SOS_OS.F:1166-1239) from the previous order's field.
//
// Round-2 structure (what round 1's k_step measured and why it changed): with two independent CTAs per SM that each
// alternate "DMMA main loop" and "recurrence epilogue", the two CTAs fall into phase (both in the main loop at half rate,
// then both in the epilogue with the tensor pipe idle; the lag between them is neutrally stable), and a concurrent epilogue
// slows the other CTA's main loop by 30 %: the pipe was 67 % busy although operand delivery alone sustains 95 %
// (tools/mainloop_bench.cu).  Here the roles are separate warps of ONE persistent CTA per SM, coupled only by mbarriers:
//
//   producer warp (1 lane)   fetches work units (item, direction, row tile) from a global counter and streams the
//                            operands of the linear (unit, level chunk, k-slab) sequence through a 4-stage shared-memory
//                            pipeline with TMA bulk copies (A slab 16 KB, rank-4 molecular slab 1 KB, field slab 8.7 KB),
//                            plus the per-chunk layer tables (XDEL, YDEL, dtau, 1/dtau, exp(-dtau/mu)); never drains.
//   8 MMA warps              16 rows x 64 levels each: nothing but DMMA over the k-slabs; at the end of a level chunk the raw
//                            accumulators go to a staging tile (swizzled 16-byte slots: conflict-free for the writer and for
//                            the row-per-thread reader) and the next chunk starts at once.
//   4 recurrence warps       one thread per tile row: source function, layer constants and the serial recurrence
//                            z <- z*a + c fused in ONE pass over the chunk's levels (blocks of 8 levels: the constants are
//                            independent, only the 8 FMAs are serial), new field stored straight to HBM (16-byte stores);
//                            the ground boundary values of a unit are formed here too, under the MMA warps' first chunk.
//
// The main loop of chunk c+1 runs while the recurrence warps work on chunk c; the tensor pipe only waits for the
// accumulator hand-off.  Work units are fetched dynamically (atomic counter), so ragged profiles (2..10 chunks) balance.
#include "sosgpu_internal.h"
#include <math.h>
#include <algorithm>
#include <cstdio>

#define SW_STG 4                       // operand pipeline stages
#define SW_MMA_WARPS 8
#define SW_EPI_WARPS 4
#define SW_THREADS ((SW_MMA_WARPS + SW_EPI_WARPS + 1) * 32)
#define SW_A_BYTES (128 * SOS_KB * 8)
#define SW_V_BYTES (8 * SOS_KB * 8)
#define SW_B_BYTES (SOS_KB * SOS_SB * 8)
#define SW_STAGE (SW_A_BYTES + SW_V_BYTES + SW_B_BYTES)
#define SW_ACC_PITCH 72                // doubles per staging-tile row (36 16-byte slots)

namespace {

__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long *bar, int count)
{
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long *bar, unsigned bytes)
{
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(unsigned long long *bar)
{
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long *bar, unsigned parity)
{
  const unsigned addr = smem_u32(bar);
  unsigned ok;
  do {
    asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
                 : "=r"(ok) : "r"(addr), "r"(parity) : "memory");
  } while (!ok);
}
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, unsigned bytes, unsigned long long *bar)
{
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory"); }
__device__ __forceinline__ void dmma8x8x4(double &c0, double &c1, double a, double b)
{
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
__device__ __forceinline__ void epi_bar() { asm volatile("bar.sync 1, %0;\n" ::"n"(SW_EPI_WARPS * 32) : "memory"); }

// 16-byte slot of (row, column pair) in the staging tile: conflict-free for the accumulator-layout writer (a quarter warp
// = 2 rows x 4 pairs) and for the row-per-thread reader (a quarter warp = 8 rows x 1 pair)
__device__ __forceinline__ int acc_slot(int row, int cpair) { return row * (SW_ACC_PITCH / 2) + (cpair ^ ((row >> 1) & 3)); }

// DMMAs of one k-slab for one warp: warp tile 16 rows x (8*NFR) levels (2 x NFR fragments of 8x8, K = 16 in 4 steps).
// NFR = column blocks (8 levels) of the chunk that hold real levels; KS = k-steps of the slab that hold real rows
// (the zero padding at the end of a direction block is skipped).
template <int NFR, int LR, int OWN, int KS>
__device__ __forceinline__ void slab_mma(double (&acc)[2][8][2], double (&tacc)[2], const double *__restrict__ a,
                                         const double *__restrict__ v, const double *__restrict__ b, int wr, int gq, int tq)
{
  const int swz = 4 * (gq & 3);
#pragma unroll
  for (int ks4 = 0; ks4 < KS; ++ks4) {
    const int kc = (ks4 * 4 + tq) ^ swz;
    if (OWN) {
      const double a0 = a[kc], a1 = a[8 * SOS_KB + kc];
      double bv[NFR];
#pragma unroll
      for (int ni = 0; ni < NFR; ++ni) bv[ni] = b[ks4 * 4 * SOS_SB + ni * 8];
#pragma unroll
      for (int ni = 0; ni < NFR; ++ni) {
        dmma8x8x4(acc[0][ni][0], acc[0][ni][1], a0, bv[ni]);
        dmma8x8x4(acc[1][ni][0], acc[1][ni][1], a1, bv[ni]);
      }
    }
    if (LR && wr < NFR) dmma8x8x4(tacc[0], tacc[1], v[kc], b[ks4 * 4 * SOS_SB + wr * 8]);   // T = V X, column block wr
  }
}

__device__ __forceinline__ bool mbar_test(unsigned long long *bar, unsigned parity)
{
  unsigned ok;
  asm volatile("{\n.reg .pred p;\nmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
               : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
  return ok != 0;
}

// The k-slabs of one level chunk for one warp.  The readiness of the NEXT stage is probed while the DMMAs of the current
// slab are in flight, so the blocking wait at the top of a slab is normally skipped.
template <int NFR, int LR, int OWN>
__device__ __forceinline__ void chunk_mma(double (&acc)[2][8][2], double (&tacc)[2], const unsigned char *sStage,
                                          unsigned long long *full, unsigned long long *empty, unsigned &q, int n_slab, int HB,
                                          int N3, int wr, int lane, int gq, int tq)
{
  bool ready = mbar_test(full + q % SW_STG, (q / SW_STG) & 1);
  for (int slab = 0; slab < n_slab; ++slab, ++q) {
    const int st = q % SW_STG;
    if (!ready) mbar_wait(full + st, (q / SW_STG) & 1);
    const unsigned char *sp = sStage + st * SW_STAGE;
    const double *a = reinterpret_cast<const double *>(sp) + (wr * 16 + gq) * SOS_KB;
    const double *v = reinterpret_cast<const double *>(sp + SW_A_BYTES) + gq * SOS_KB;
    const double *b = reinterpret_cast<const double *>(sp + SW_A_BYTES + SW_V_BYTES) + tq * SOS_SB + gq;
    const int kq = (slab * SOS_KB) % HB;
    ready = mbar_test(full + (q + 1) % SW_STG, ((q + 1) / SW_STG) & 1);   // result is consumed after this slab's DMMAs
    if (kq + SOS_KB <= N3) slab_mma<NFR, LR, OWN, 4>(acc, tacc, a, v, b, wr, gq, tq);
    else {                                                     // tail of a direction block: k-steps of pure padding skipped
      const int ks_lim = max(0, (N3 - kq + 3) >> 2);
      if (ks_lim == 3) slab_mma<NFR, LR, OWN, 3>(acc, tacc, a, v, b, wr, gq, tq);
      else if (ks_lim == 2) slab_mma<NFR, LR, OWN, 2>(acc, tacc, a, v, b, wr, gq, tq);
      else if (ks_lim == 1) slab_mma<NFR, LR, OWN, 1>(acc, tacc, a, v, b, wr, gq, tq);
      else if (ks_lim >= 4) slab_mma<NFR, LR, OWN, 4>(acc, tacc, a, v, b, wr, gq, tq);
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(empty + st);
  }
}

template <int LR, int OWN>
__device__ __forceinline__ void chunk_dispatch(int nfr, double (&acc)[2][8][2], double (&tacc)[2], const unsigned char *sStage,
                                               unsigned long long *full, unsigned long long *empty, unsigned &q, int n_slab, int HB,
                                               int N3, int wr, int lane, int gq, int tq)
{
  switch (nfr) {
    case 8: chunk_mma<8, LR, OWN>(acc, tacc, sStage, full, empty, q, n_slab, HB, N3, wr, lane, gq, tq); break;
    case 7: chunk_mma<7, LR, OWN>(acc, tacc, sStage, full, empty, q, n_slab, HB, N3, wr, lane, gq, tq); break;
    case 6: chunk_mma<6, LR, OWN>(acc, tacc, sStage, full, empty, q, n_slab, HB, N3, wr, lane, gq, tq); break;
    case 5: chunk_mma<5, LR, OWN>(acc, tacc, sStage, full, empty, q, n_slab, HB, N3, wr, lane, gq, tq); break;
    case 4: chunk_mma<4, LR, OWN>(acc, tacc, sStage, full, empty, q, n_slab, HB, N3, wr, lane, gq, tq); break;
    case 3: chunk_mma<3, LR, OWN>(acc, tacc, sStage, full, empty, q, n_slab, HB, N3, wr, lane, gq, tq); break;
    case 2: chunk_mma<2, LR, OWN>(acc, tacc, sStage, full, empty, q, n_slab, HB, N3, wr, lane, gq, tq); break;
    default: chunk_mma<1, LR, OWN>(acc, tacc, sStage, full, empty, q, n_slab, HB, N3, wr, lane, gq, tq); break;
  }
}

struct Unit {                                                  // decoded work unit (uniform per CTA), 64 bytes
  int item, dir, g0, ng, R, r0, N, HB, KP, NT, L, n_chunk, n_slab, lr, valid, w;
};

__device__ __forceinline__ Unit decode_unit(int w, const ItemDev *items, const TermDev *terms, const OpticsDev *optics,
                                            const KsetDev *ksets, const int *list, int tiles_per_dir)
{
  Unit u;
  u.w = w;
  const int per_item = 2 * tiles_per_dir;
  const int ii = w / per_item, t = w - ii * per_item;
  u.item = list ? list[ii] : ii;
  const ItemDev &it = items[u.item];
  const TermDev &tm = terms[it.term];
  const OpticsDev &op = optics[tm.optics];
  u.N = op.nbmu; u.HB = op.HB; u.KP = op.KP; u.NT = tm.nt; u.L = tm.nt + 1;
  u.dir = t / tiles_per_dir;
  const int tile = t - u.dir * tiles_per_dir;
  const int groups = u.HB >> 4;
  const int gpt = (groups + tiles_per_dir - 1) / tiles_per_dir;
  u.g0 = tile * gpt;
  u.valid = (u.g0 < groups) && it.active;                      // inactive: IGMAX < 2 (SOS_OS.F:1152)
  u.ng = u.valid ? min(gpt, groups - u.g0) : 0;
  u.R = u.ng * 16;
  u.r0 = u.dir * u.HB + u.g0 * 16;
  u.n_chunk = u.valid ? (u.L + SOS_CH - 1) / SOS_CH : 0;
  u.n_slab = u.KP / SOS_KB;
  u.lr = ksets[it.kset].dual;
  return u;
}

}  // namespace

__global__ void __launch_bounds__(SW_THREADS, 1)
k_sweep(const ItemDev *__restrict__ items, const TermDev *__restrict__ terms, const OpticsDev *__restrict__ optics,
        const KsetDev *__restrict__ ksets, const int *__restrict__ list, const int *__restrict__ count_ptr, int count_fixed,
        int tiles_per_dir, int att_rows_cap, unsigned *__restrict__ work_counter, double *__restrict__ jdump)
{
  extern __shared__ __align__(128) unsigned char smem_raw[];
  // ---- shared memory carve ----
  unsigned char *sStage = smem_raw;
  double *sAcc = reinterpret_cast<double *>(smem_raw + SW_STG * SW_STAGE);          // [128][72] swizzled slots
  double *sT = sAcc + 128 * SW_ACC_PITCH;                                           // [4][64] raw molecular functionals
  double *sXd = sT + 4 * SOS_CH;                                                    // [<=66] XDEL of the chunk's levels
  double *sYd = sXd + 72;
  double *sDt = sYd + 72;                                                           // [<=66] layers lb_al..
  double *sInv = sDt + 72;
  double *sG = sInv + 72;                                                           // [3*80] ground values of the downward field
  unsigned long long *bars = reinterpret_cast<unsigned long long *>(sG + 3 * 80);
  unsigned long long *full = bars, *empty = bars + SW_STG;
  unsigned long long *acc_full = bars + 2 * SW_STG, *acc_empty = acc_full + 1, *tab_full = acc_full + 2, *tab_empty = acc_full + 3;
  unsigned long long *work_full = acc_full + 4, *work_empty = acc_full + 6;       // [2] each
  Unit *sUnit = reinterpret_cast<Unit *>(acc_full + 8);                             // [2] decoded work units (64 bytes each)
  double *sAtt = reinterpret_cast<double *>(acc_full + 24);                         // [<=66][N] exp(-dtau/mu_k)

  const int tid = threadIdx.x, lane = tid & 31, wr = tid >> 5;
  if (tid == 0) {
    for (int s = 0; s < SW_STG; ++s) { mbar_init(full + s, 1); mbar_init(empty + s, SW_MMA_WARPS); }
    mbar_init(acc_full, SW_MMA_WARPS); mbar_init(acc_empty, SW_EPI_WARPS);
    mbar_init(tab_full, 1); mbar_init(tab_empty, SW_EPI_WARPS);
    for (int s = 0; s < 2; ++s) { mbar_init(work_full + s, 1); mbar_init(work_empty + s, SW_MMA_WARPS + SW_EPI_WARPS); }
    fence_proxy_async();
  }
  __syncthreads();
  const int nwork = (count_ptr ? *count_ptr : count_fixed) * 2 * tiles_per_dir;

  if (wr == SW_MMA_WARPS + SW_EPI_WARPS) {
    // =============================== producer warp ===============================
    if (lane != 0) return;
    unsigned q = 0, tabn = 0, wn = 0;
    while (true) {
      int w = (int)atomicAdd(work_counter, 1u);
      if (w >= nwork) w = -1;
      const int slot = wn & 1;
      Unit u;
      if (w >= 0) u = decode_unit(w, items, terms, optics, ksets, list, tiles_per_dir);
      else { u = Unit{}; u.w = -1; }
      if (wn >= 2) mbar_wait(work_empty + slot, ((wn >> 1) - 1) & 1);
      sUnit[slot] = u;                                           // the other roles take the decoded unit from shared memory
      mbar_arrive(work_full + slot);
      ++wn;
      if (w < 0) break;
      if (!u.valid) continue;
      const ItemDev &it = items[u.item];
      const TermDev &tm = terms[it.term];
      const KsetDev &ks = ksets[it.kset];
      const double *xprev = it.x[it.n & 1];
      const double *Ag = ks.apackA + (size_t)u.r0 * SOS_KB;
      const double *Vg = u.lr ? ks.vpack + (size_t)u.dir * 8 * u.KP : nullptr;
      const unsigned tx = (unsigned)(u.R * SOS_KB * 8 + (u.lr ? SW_V_BYTES : 0) + SW_B_BYTES);
      const bool up = (u.dir == 0);
      for (int chunk = 0; chunk < u.n_chunk; ++chunk) {
        const int ci = up ? (u.n_chunk - 1 - chunk) : chunk;
        const int c0 = ci * SOS_CH;
        for (int slab = 0; slab < u.n_slab; ++slab, ++q) {
          const int st = q % SW_STG;
          if (q >= SW_STG) mbar_wait(empty + st, ((q / SW_STG) - 1) & 1);
          unsigned char *sp = sStage + st * SW_STAGE;
          mbar_expect_tx(full + st, tx);
          bulk_g2s(sp, Ag + (size_t)slab * u.KP * SOS_KB, (unsigned)(u.R * SOS_KB * 8), full + st);
          if (u.lr) bulk_g2s(sp + SW_A_BYTES, Vg + (size_t)slab * 8 * SOS_KB, SW_V_BYTES, full + st);
          bulk_g2s(sp + SW_A_BYTES + SW_V_BYTES, xprev + SOS_XIDX(u.KP, slab * SOS_KB, c0), SW_B_BYTES, full + st);
        }
        // layer tables of this chunk (needed once its main loop is over): rows [lb_al, le], even start and count
        const int lb_al = max(c0 - 1, 0) & ~1;
        const int le = min(c0 + SOS_CH - 1, u.NT - 1);
        const int nrow = (le - lb_al + 2) & ~1;
        const int nlev = (min(SOS_CH, u.L - c0) + 1) & ~1;
        const bool att_staged = nrow <= att_rows_cap;
        if (tabn >= 1) mbar_wait(tab_empty, (tabn - 1) & 1);
        mbar_expect_tx(tab_full, (unsigned)(nrow * 16 + nlev * 16 + (att_staged ? nrow * u.N * 8 : 0)));
        bulk_g2s(sXd, tm.xdel + c0, (unsigned)(nlev * 8), tab_full);
        bulk_g2s(sYd, tm.ydel + c0, (unsigned)(nlev * 8), tab_full);
        bulk_g2s(sDt, tm.dt + lb_al, (unsigned)(nrow * 8), tab_full);
        bulk_g2s(sInv, tm.inv + lb_al, (unsigned)(nrow * 8), tab_full);
        if (att_staged) bulk_g2s(sAtt, tm.att + (size_t)lb_al * u.N, (unsigned)(nrow * u.N * 8), tab_full);
        ++tabn;
      }
    }
    return;
  }

  if (wr < SW_MMA_WARPS) {
    // =============================== MMA warps ===============================
    const int gq = lane >> 2, tq = lane & 3;
    unsigned q = 0, accn = 0, wn = 0;
    double acc[2][8][2];
    double tacc[2];
    while (true) {
      const int slot = wn & 1;
      mbar_wait(work_full + slot, (wn >> 1) & 1);
      const Unit u = sUnit[slot];
      __syncwarp();
      if (lane == 0) mbar_arrive(work_empty + slot);
      ++wn;
      if (u.w < 0) break;
      if (!u.valid) continue;
      const bool own = wr < u.ng;
      const bool up = (u.dir == 0);
      for (int chunk = 0; chunk < u.n_chunk; ++chunk) {
        const int ci = up ? (u.n_chunk - 1 - chunk) : chunk;
        const int c0 = ci * SOS_CH;
        const int nfr = (min(SOS_CH, u.L - c0) + 7) >> 3;
#pragma unroll
        for (int mi = 0; mi < 2; ++mi)
#pragma unroll
          for (int ni = 0; ni < 8; ++ni) { acc[mi][ni][0] = 0.0; acc[mi][ni][1] = 0.0; }
        tacc[0] = tacc[1] = 0.0;
        if (u.lr) {
          if (own) chunk_dispatch<1, 1>(nfr, acc, tacc, sStage, full, empty, q, u.n_slab, u.HB, 3 * u.N, wr, lane, gq, tq);
          else chunk_dispatch<1, 0>(nfr, acc, tacc, sStage, full, empty, q, u.n_slab, u.HB, 3 * u.N, wr, lane, gq, tq);
        } else {
          if (own) chunk_dispatch<0, 1>(nfr, acc, tacc, sStage, full, empty, q, u.n_slab, u.HB, 3 * u.N, wr, lane, gq, tq);
          else chunk_dispatch<0, 0>(nfr, acc, tacc, sStage, full, empty, q, u.n_slab, u.HB, 3 * u.N, wr, lane, gq, tq);
        }
        // ---- hand the raw accumulators to the recurrence warps ----
        if (accn >= 1) mbar_wait(acc_empty, (accn - 1) & 1);
        if (own) {
          double2 *sA2 = reinterpret_cast<double2 *>(sAcc);
#pragma unroll
          for (int mi = 0; mi < 2; ++mi) {
            const int row = wr * 16 + mi * 8 + gq;
#pragma unroll
            for (int ni = 0; ni < 8; ++ni) sA2[acc_slot(row, ni * 4 + tq)] = make_double2(acc[mi][ni][0], acc[mi][ni][1]);
          }
        }
        if (u.lr && gq < 4 && wr < nfr) { sT[gq * SOS_CH + wr * 8 + 2 * tq] = tacc[0]; sT[gq * SOS_CH + wr * 8 + 2 * tq + 1] = tacc[1]; }
        __syncwarp();
        if (lane == 0) mbar_arrive(acc_full);
        ++accn;
      }
    }
    return;
  }

  // =============================== recurrence warps: one thread per tile row ===============================
  {
    const int e = tid - SW_MMA_WARPS * 32;                       // tile row of this thread
    unsigned accn = 0, wn = 0;
    while (true) {
      const int slot = wn & 1;
      mbar_wait(work_full + slot, (wn >> 1) & 1);
      const Unit u = sUnit[slot];
      __syncwarp();
      if (lane == 0) mbar_arrive(work_empty + slot);
      ++wn;
      if (u.w < 0) break;
      if (!u.valid) continue;
      const ItemDev &it = items[u.item];
      const TermDev &tm = terms[it.term];
      const OpticsDev &op = optics[tm.optics];
      const KsetDev &ks = ksets[it.kset];
      const int N = u.N, NT = u.NT, KP = u.KP, HB = u.HB, L = u.L;
      const bool up = (u.dir == 0);
      const double *__restrict__ xprev = it.x[it.n & 1];
      double *__restrict__ xnext = it.x[(it.n + 1) & 1];
      const int qrow = (e < u.R) ? (u.g0 * 16 + e) : 3 * N;      // row within the direction block
      const bool rowvalid = qrow < 3 * N;
      const int so = rowvalid ? qrow / N : 0, kk = rowvalid ? qrow % N + 1 : 1;
      const double mu = op.rmu[kk + N];
      const double us = (u.lr && rowvalid) ? ks.urow[u.r0 + e] : 0.0;
      const double u0 = (so == 0 && rowvalid) ? 1.0 : 0.0;
      // ---- ground boundary value of this row (mu > 0 rows; SOS_OS.F:1166-1239) ----
      double bc = 0.0;
      if (up) {
        for (int c = e; c < 3 * N; c += SW_EPI_WARPS * 32) sG[c] = xprev[SOS_XIDX(KP, HB + c, NT)];
        epi_bar();
        if (rowvalid) {
          double xr = 0.0;
          if (!(op.ro == 0.0 || it.is != 0)) {                   // Lambert (SOS_OS.F:1177-1190)
            double lsol = 0.0;
            for (int j = 1; j <= N; ++j) lsol = lsol + op.ga[j + N] * sG[j - 1] * op.rmu[j + N];
            lsol = 2 * lsol * op.ro;
            xr = lsol;
            if (so == 0) bc = lsol;
          }
          if (op.imat_surf == 1) {                               // BRDF / BPDF quadrature (SOS_OS.F:1194-1220)
            const float *rs = op.surf + (size_t)it.is * 9 * N * N + (size_t)(so * 3) * N * N + (size_t)(kk - 1) * N;
            const bool pol = (op.ipolar != 0);
            double accs = 0.0;
            for (int j = 1; j <= N; ++j) {
              double r1 = (double)rs[j - 1], r2 = (double)rs[(size_t)N * N + j - 1], r3 = (double)rs[(size_t)2 * N * N + j - 1];
              if (!pol) { if (so == 0) { r2 = 0.0; r3 = 0.0; } else { r1 = 0.0; r2 = 0.0; r3 = 0.0; } }
              accs = accs + op.ga[j + N] * (sG[j - 1] * r1 + sG[N + j - 1] * r2 + sG[2 * N + j - 1] * r3);
            }
            const double rrmu = 2 / mu;
            bc = (so == 0) ? accs * rrmu + xr : accs * rrmu;
          }
          if (op.ifresnel == 1) {                                // flat sea (SOS_OS.F:1225-1239)
            const double gi = sG[kk - 1], gqv = sG[N + kk - 1], gu = sG[2 * N + kk - 1];
            if (so == 0) bc = bc + op.f11[kk] * gi + op.f12[kk] * gqv;
            else if (so == 1) bc = bc + op.f12[kk] * gi + op.f11[kk] * gqv;
            else bc = bc + op.f33[kk] * gu;
          }
        }
        epi_bar();                                               // sG may be overwritten by the next unit
      }

      const double2 *sA2 = reinterpret_cast<const double2 *>(sAcc);
      const int tsel = (so + 1) * SOS_CH;                        // sT row of this row's Stokes type
      double z = 0.0;                                            // recurrence state
      double sedge = 0.0;                                        // source function at the neighbouring level of the previous chunk
      const double rmuk = -mu;
      for (int chunk = 0; chunk < u.n_chunk; ++chunk) {
        const int ci = up ? (u.n_chunk - 1 - chunk) : chunk;
        const int c0 = ci * SOS_CH;
        const int lb_al = max(c0 - 1, 0) & ~1;
        const int le = min(c0 + SOS_CH - 1, NT - 1);
        const int nrow = (le - lb_al + 2) & ~1;
        const bool att_staged = nrow <= att_rows_cap;
        mbar_wait(tab_full, accn & 1);
        mbar_wait(acc_full, accn & 1);
        if (rowvalid) {
          // layer l of this row: attenuation a(l), dtau(l), 1/dtau(l)
          const double *attp = att_staged ? (sAtt + (kk - 1) - (size_t)lb_al * N) : (tm.att + (kk - 1));
          const double *dtp = sDt - lb_al, *ivp = sInv - lb_al;
          const int top = min(c0 + SOS_CH - 1, NT);              // highest real level of the chunk
          double *xrow = xnext + SOS_XIDX(KP, u.r0 + e, c0);     // element (row, level c0)
          double *jrow = jdump ? jdump + (size_t)(u.r0 + e) * tm.LP + c0 : nullptr;
          // source function of column col from the raw accumulator (SOS_FSOURCE_ORDREIG, molecular part in factored form)
          auto source = [&](double araw, int col) -> double {
            double v = sXd[col] * araw;
            if (u.lr) v = v + sYd[col] * (u0 * sT[col] + us * sT[tsel + col]);
            return v;
          };
          auto raw = [&](int col) -> double {
            const double2 t = sA2[acc_slot(e, col >> 1)];
            return (col & 1) ? t.y : t.x;
          };
          if (up) {
            // ---- from the ground upwards: levels top .. c0 (SOS_OS.F:2279-2310) ----
            int lv = top;
            const int full_hi = jdump ? c0 - 1 : ((top == NT) ? (NT & ~7) - 1 : top);   // levels c0 .. full_hi: whole blocks of 8 below NT
            for (; lv > full_hi; --lv) {                         // ragged head (at most 8 levels; holds level NT)
              const int col = lv - c0;
              const double s = source(raw(col), col);
              if (lv == NT) z = bc;
              else {
                const double a = attp[(size_t)lv * N], dl = dtp[lv], iv = ivp[lv];
                const double A = (sedge - s) * iv;
                z = z * a + ((1.0 - a) * (A * mu + s) - A * (a * dl));
              }
              sedge = s;
              xrow[col] = z;
              if (jrow) jrow[col] = s;
            }
            for (; lv >= c0; lv -= 8) {                          // whole blocks of 8 levels, lv = highest level of the block
              const int cb = lv - 7 - c0;                        // first column of the block (multiple of 8)
              double S[9], cst[8], aa[8];
#pragma unroll
              for (int p = 0; p < 4; ++p) {
                const double2 t = sA2[acc_slot(e, (cb >> 1) + p)];
                S[2 * p] = source(t.x, cb + 2 * p);
                S[2 * p + 1] = source(t.y, cb + 2 * p + 1);
              }
              S[8] = sedge;
#pragma unroll
              for (int j = 0; j < 8; ++j) {
                const int l = c0 + cb + j;
                const double a = attp[(size_t)l * N], dl = dtp[l], iv = ivp[l];
                const double A = (S[j + 1] - S[j]) * iv;
                cst[j] = (1.0 - a) * (A * mu + S[j]) - A * (a * dl);
                aa[j] = a;
              }
              double o[8];
#pragma unroll
              for (int j = 7; j >= 0; --j) { z = z * aa[j] + cst[j]; o[j] = z; }
              sedge = S[0];
#pragma unroll
              for (int p = 0; p < 4; ++p) *reinterpret_cast<double2 *>(xrow + cb + 2 * p) = make_double2(o[2 * p], o[2 * p + 1]);
            }
          } else {
            // ---- from the top downwards: levels c0 .. top (SOS_OS.F:2320-2354) ----
            int lv = c0;
            const int n_full = jdump ? 0 : ((top + 1 - c0) >> 3);  // whole blocks of 8 in this chunk
            for (int blk = 0; blk < n_full; ++blk, lv += 8) {
              const int cb = lv - c0;
              double S[9], cst[8], aa[8];
              S[0] = sedge;
#pragma unroll
              for (int p = 0; p < 4; ++p) {
                const double2 t = sA2[acc_slot(e, (cb >> 1) + p)];
                S[2 * p + 1] = source(t.x, cb + 2 * p);
                S[2 * p + 2] = source(t.y, cb + 2 * p + 1);
              }
#pragma unroll
              for (int j = 0; j < 8; ++j) {
                const int l = max(lv + j - 1, 0);                // layer above level lv+j (level 0 has none: fixed below)
                const double a = attp[(size_t)l * N], dl = dtp[l], iv = ivp[l];
                const double A = (S[j + 1] - S[j]) * iv;
                cst[j] = (1.0 - a) * (A * rmuk + S[j + 1]) + A * (a * dl);
                aa[j] = a;
              }
              if (lv == 0) { cst[0] = 0.0; aa[0] = 0.0; }         // level 0: X = 0
              double o[8];
#pragma unroll
              for (int j = 0; j < 8; ++j) { z = z * aa[j] + cst[j]; o[j] = z; }
              sedge = S[8];
#pragma unroll
              for (int p = 0; p < 4; ++p) *reinterpret_cast<double2 *>(xrow + cb + 2 * p) = make_double2(o[2 * p], o[2 * p + 1]);
            }
            for (; lv <= top; ++lv) {                            // ragged tail
              const int col = lv - c0;
              const double s = source(raw(col), col);
              if (lv == 0) z = 0.0;
              else {
                const double a = attp[(size_t)(lv - 1) * N], dl = dtp[lv - 1], iv = ivp[lv - 1];
                const double A = (s - sedge) * iv;
                z = z * a + ((1.0 - a) * (A * rmuk + s) + A * (a * dl));
              }
              sedge = s;
              xrow[col] = z;
              if (jrow) jrow[col] = s;
            }
          }
        }
        __syncwarp();
        if (lane == 0) { mbar_arrive(acc_empty); mbar_arrive(tab_empty); }
        ++accn;
      }
    }
  }
}

static size_t sweep_smem_bytes(int att_rows_cap, int nmax)
{
  return (size_t)SW_STG * SW_STAGE + (size_t)128 * SW_ACC_PITCH * 8 + (4 * SOS_CH + 4 * 72 + 3 * 80) * 8 + (2 * SW_STG + 24) * 8 +
         (size_t)att_rows_cap * nmax * 8 + 128;
}

// One scattering order for the items of `list` (or 0..nitem-1).  The number of items comes from device memory when
// count_ptr != null (the wave loop never waits for it); nitem is then only an upper bound.
extern "C" int sos_launch_sweep(const ItemDev *items, const TermDev *terms, const OpticsDev *optics, const KsetDev *ksets,
                                const int *list, const int *count_ptr, int nitem, int maxHB, unsigned *work_counter, int num_sms,
                                double *jdump, cudaStream_t st)
{
  if (nitem <= 0) return 0;
  const int groups = maxHB / 16;
  const int tiles_per_dir = (groups + SW_MMA_WARPS - 1) / SW_MMA_WARPS;
  const int nmax = maxHB / 3;                                    // 3N <= HB
  int att_rows = 66;
  while (att_rows > 0 && sweep_smem_bytes(att_rows, nmax) > 227 * 1024) att_rows = 0;   // all or nothing
  const size_t smem = sweep_smem_bytes(att_rows, nmax);
  cudaFuncSetAttribute(k_sweep, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  cudaMemsetAsync(work_counter, 0, sizeof(unsigned), st);
  const long long units = (long long)nitem * 2 * tiles_per_dir;
  const int grid = (int)std::min<long long>(num_sms, units);
  k_sweep<<<grid, SW_THREADS, smem, st>>>(items, terms, optics, ksets, list, count_ptr, nitem, tiles_per_dir, att_rows, work_counter, jdump);
  return 1;
}
