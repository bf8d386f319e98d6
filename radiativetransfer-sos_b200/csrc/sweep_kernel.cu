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
#include "sosgpu_async.cuh"
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
                                         const double *__restrict__ v, const double *__restrict__ b, const double *__restrict__ bt,
                                         bool do_t, int gq, int tq)
{
  const int swz = 4 * (gq & 3);
#pragma unroll
  for (int ks4 = 0; ks4 < KS; ++ks4) {
    const int kc = (ks4 * 4 + tq) ^ swz;
    if (OWN && NFR > 0) {
      const double a0 = a[kc], a1 = a[8 * SOS_KB + kc];
      double bv[NFR > 0 ? NFR : 1];
#pragma unroll
      for (int ni = 0; ni < NFR; ++ni) bv[ni] = b[ks4 * 4 * SOS_SB + ni * 8];
#pragma unroll
      for (int ni = 0; ni < NFR; ++ni) {
        dmma8x8x4(acc[0][ni][0], acc[0][ni][1], a0, bv[ni]);
        dmma8x8x4(acc[1][ni][0], acc[1][ni][1], a1, bv[ni]);
      }
    }
    if (LR && do_t) dmma8x8x4(tacc[0], tacc[1], v[kc], bt[ks4 * 4 * SOS_SB]);   // T = V X, this warp's column block
  }
}


// The k-slabs of one level chunk for one warp.  The readiness of the NEXT stage is probed while the DMMAs of the current
// slab are in flight, so the blocking wait at the top of a slab is normally skipped.
template <int NFR, int LR, int OWN>
__device__ __forceinline__ void chunk_mma(double (&acc)[2][8][2], double (&tacc)[2], const unsigned char *sStage,
                                          unsigned long long *full, unsigned long long *empty, unsigned &q, int n_slab, int HB,
                                          int N3, int rg, int col0, int tcol, int lane, int gq, int tq)
{
  const bool do_t = tcol >= 0;                                   // this warp owns column block tcol / 8 of T = V X
  bool ready = mbar_test(full + q % SW_STG, (q / SW_STG) & 1);
  int kq = 0;                                                  // first k-row of the slab within its direction block
  for (int slab = 0; slab < n_slab; ++slab, ++q) {
    const int st = q % SW_STG;
    if (!ready) mbar_wait(full + st, (q / SW_STG) & 1);
    const unsigned char *sp = sStage + st * SW_STAGE;
    const double *a = reinterpret_cast<const double *>(sp) + (rg * 16 + gq) * SOS_KB;
    const double *v = reinterpret_cast<const double *>(sp + SW_A_BYTES) + gq * SOS_KB;
    const double *b0 = reinterpret_cast<const double *>(sp + SW_A_BYTES + SW_V_BYTES) + tq * SOS_SB + gq;
    const double *b = b0 + col0, *bt = b0 + (do_t ? tcol : 0);
    ready = mbar_test(full + (q + 1) % SW_STG, ((q + 1) / SW_STG) & 1);   // result is consumed after this slab's DMMAs
    if (kq + SOS_KB <= N3) slab_mma<NFR, LR, OWN, 4>(acc, tacc, a, v, b, bt, do_t, gq, tq);
    else {                                                     // tail of a direction block: k-steps of pure padding skipped
      const int ks_lim = max(0, (N3 - kq + 3) >> 2);
      if (ks_lim == 3) slab_mma<NFR, LR, OWN, 3>(acc, tacc, a, v, b, bt, do_t, gq, tq);
      else if (ks_lim == 2) slab_mma<NFR, LR, OWN, 2>(acc, tacc, a, v, b, bt, do_t, gq, tq);
      else if (ks_lim == 1) slab_mma<NFR, LR, OWN, 1>(acc, tacc, a, v, b, bt, do_t, gq, tq);
      else if (ks_lim >= 4) slab_mma<NFR, LR, OWN, 4>(acc, tacc, a, v, b, bt, do_t, gq, tq);
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(empty + st);
    kq += SOS_KB;
    if (kq >= HB) kq -= HB;
  }
}

template <int LR, int OWN>
__device__ __forceinline__ void chunk_dispatch(int live, double (&acc)[2][8][2], double (&tacc)[2], const unsigned char *sStage,
                                               unsigned long long *full, unsigned long long *empty, unsigned &q, int n_slab, int HB,
                                               int N3, int rg, int col0, int tcol, int lane, int gq, int tq)
{
#define SW_CASE(n) case n: chunk_mma<n, LR, OWN>(acc, tacc, sStage, full, empty, q, n_slab, HB, N3, rg, col0, tcol, lane, gq, tq); break;
  switch (live) {                                              // live 8-level column blocks of this warp in this chunk
    SW_CASE(8) SW_CASE(7) SW_CASE(6) SW_CASE(5) SW_CASE(4) SW_CASE(3) SW_CASE(2) SW_CASE(1)
    default: chunk_mma<0, LR, OWN>(acc, tacc, sStage, full, empty, q, n_slab, HB, N3, rg, col0, tcol, lane, gq, tq); break;
  }
#undef SW_CASE
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

// MW = number of DMMA warps: 8 (16 rows x 64 levels each, the shipped instantiation) or 16 (16 rows x 32 levels each: warp w
// owns row group w & 7 and column half w >> 3 of the 128 x 64 tile; measured slower, not instantiated)
template <int MW>
__global__ void __launch_bounds__((MW + SW_EPI_WARPS + 1) * 32, 1)   // warps are allocated in fours: 13 -> 16 (128 regs), 21 -> 24 (80 regs)
k_sweep(const ItemDev *__restrict__ items, const TermDev *__restrict__ terms, const OpticsDev *__restrict__ optics,
        const KsetDev *__restrict__ ksets, const int *__restrict__ list, const int *__restrict__ count_ptr, int count_fixed,
        int tiles_per_dir, unsigned *__restrict__ work_counter, double *__restrict__ jdump)
{
  extern __shared__ __align__(128) unsigned char smem_raw[];
  // ---- shared memory carve ----
  unsigned char *sStage = smem_raw;
  double *sAcc = reinterpret_cast<double *>(smem_raw + SW_STG * SW_STAGE);          // [128][72] swizzled slots
  double *sT = sAcc + 128 * SW_ACC_PITCH;                                           // [4][64] raw molecular functionals
  double *sXd = sT + 4 * SOS_CH;                                                    // [<=66] XDEL of the chunk's levels
  double *sYd = sXd + 72;
  double *sG = sYd + 72;                                                           // [3*80] ground values of the downward field
  unsigned long long *bars = reinterpret_cast<unsigned long long *>(sG + 3 * 80);
  unsigned long long *full = bars, *empty = bars + SW_STG;
  unsigned long long *acc_full = bars + 2 * SW_STG, *acc_empty = acc_full + 1, *tab_full = acc_full + 2, *tab_empty = acc_full + 3;
  unsigned long long *work_full = acc_full + 4, *work_empty = acc_full + 6;       // [2] each
  Unit *sUnit = reinterpret_cast<Unit *>(acc_full + 8);                             // [2] decoded work units (64 bytes each)

  const int tid = threadIdx.x, lane = tid & 31, wr = tid >> 5;
  if (tid == 0) {
    for (int s = 0; s < SW_STG; ++s) { mbar_init(full + s, 1); mbar_init(empty + s, MW); }
    mbar_init(acc_full, MW); mbar_init(acc_empty, SW_EPI_WARPS);
    mbar_init(tab_full, 1); mbar_init(tab_empty, SW_EPI_WARPS);
    for (int s = 0; s < 2; ++s) { mbar_init(work_full + s, 1); mbar_init(work_empty + s, MW + SW_EPI_WARPS); }
    fence_proxy_async();
  }
  __syncthreads();
  const int nwork = (count_ptr ? *count_ptr : count_fixed) * 2 * tiles_per_dir;

  if (wr == MW + SW_EPI_WARPS) {
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
        // XDEL / YDEL of this chunk's levels (needed once its main loop is over)
        const int nlev = (min(SOS_CH, u.L - c0) + 1) & ~1;
        if (tabn >= 1) mbar_wait(tab_empty, (tabn - 1) & 1);
        mbar_expect_tx(tab_full, (unsigned)(nlev * 16));
        bulk_g2s(sXd, tm.xdel + c0, (unsigned)(nlev * 8), tab_full);
        bulk_g2s(sYd, tm.ydel + c0, (unsigned)(nlev * 8), tab_full);
        ++tabn;
      }
    }
    return;
  }

  if (wr < MW) {
    // =============================== MMA warps ===============================
    constexpr int NFW = 64 / MW;                                // 8-level column blocks per warp
    const int gq = lane >> 2, tq = lane & 3;
    const int rg = wr & 7, chh = wr >> 3;                       // row group, column half
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
      const bool own = rg < u.ng;
      const bool up = (u.dir == 0);
      for (int chunk = 0; chunk < u.n_chunk; ++chunk) {
        const int ci = up ? (u.n_chunk - 1 - chunk) : chunk;
        const int c0 = ci * SOS_CH;
        const int nfr = (min(SOS_CH, u.L - c0) + 7) >> 3;
        const int live = min(NFW, max(0, nfr - chh * NFW));      // live column blocks of this warp
        const int col0 = chh * NFW * 8;
        const int tcol = (u.lr && chh == 0 && rg < nfr) ? rg * 8 : -1;   // T = V X: column block rg, by the warps of half 0
#pragma unroll
        for (int mi = 0; mi < 2; ++mi)
#pragma unroll
          for (int ni = 0; ni < 8; ++ni) { acc[mi][ni][0] = 0.0; acc[mi][ni][1] = 0.0; }
        tacc[0] = tacc[1] = 0.0;
        if (u.lr) {
          if (own) chunk_dispatch<1, 1>(live, acc, tacc, sStage, full, empty, q, u.n_slab, u.HB, 3 * u.N, rg, col0, tcol, lane, gq, tq);
          else chunk_dispatch<1, 0>(live, acc, tacc, sStage, full, empty, q, u.n_slab, u.HB, 3 * u.N, rg, col0, tcol, lane, gq, tq);
        } else {
          if (own) chunk_dispatch<0, 1>(live, acc, tacc, sStage, full, empty, q, u.n_slab, u.HB, 3 * u.N, rg, col0, tcol, lane, gq, tq);
          else chunk_dispatch<0, 0>(live, acc, tacc, sStage, full, empty, q, u.n_slab, u.HB, 3 * u.N, rg, col0, tcol, lane, gq, tq);
        }
        // ---- hand the raw accumulators to the recurrence warps ----
        if (accn >= 1) mbar_wait(acc_empty, (accn - 1) & 1);
        if (own) {
          double2 *sA2 = reinterpret_cast<double2 *>(sAcc);
#pragma unroll
          for (int mi = 0; mi < 2; ++mi) {
            const int row = rg * 16 + mi * 8 + gq;
#pragma unroll
            for (int ni = 0; ni < NFW; ++ni) sA2[acc_slot(row, (chh * NFW + ni) * 4 + tq)] = make_double2(acc[mi][ni][0], acc[mi][ni][1]);
          }
        }
        if (tcol >= 0 && gq < 4) { sT[gq * SOS_CH + tcol + 2 * tq] = tacc[0]; sT[gq * SOS_CH + tcol + 2 * tq + 1] = tacc[1]; }
        __syncwarp();
        if (lane == 0) mbar_arrive(acc_full);
        ++accn;
      }
    }
    return;
  }

  // =============================== recurrence warps: one thread per tile row ===============================
  {
    constexpr int EB = (MW == 16) ? 4 : 8;                       // levels per block of the fused pass (register budget: 80 / 128)
    const int e = tid - MW * 32;                                 // tile row of this thread
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
      double2 *sAw2 = reinterpret_cast<double2 *>(sAcc);          // the new field replaces the accumulators in place
      double *sAw = sAcc;
      const int tsel = (so + 1) * SOS_CH;                        // sT row of this row's Stokes type
      double z = 0.0;                                            // recurrence state
      double sedge = 0.0;                                        // source function at the neighbouring level of the previous chunk
      double redge = 0.0;                                        // ... and its raw accumulator (aerosol-only orders)
      for (int chunk = 0; chunk < u.n_chunk; ++chunk) {
        const int ci = up ? (u.n_chunk - 1 - chunk) : chunk;
        const int c0 = ci * SOS_CH;
        mbar_wait(tab_full, accn & 1);
        mbar_wait(acc_full, accn & 1);
        if (rowvalid) {
          // Layer l of this row: a = exp(-dtau/mu), g = (1-a)*mu/dtau - a, b = 1 - a - g (tables of k_att, coalesced over the
          // rows of a warp).  With S the source function at the two levels of a layer, SOS_INTEGR_EPOPT's update
          //   z <- z*a + (1-a)*(A*mu + S_i) -/+ A*a*dtau,  A = dS/dtau   (SOS_OS.F:2288-2309, 2332-2353)
          // is z <- z*a + b*S_i + g*S_j: 4 FP64 instructions per level.  (FP64 vector instructions share the pipe with DMMA
          // and each costs about a DMMA slot -- tools/mainloop_bench.cu -- so the recurrence warps must be frugal.)
          const double *__restrict__ attp = tm.att + (kk - 1);
          const double *__restrict__ gp = tm.gco + (kk - 1);
          const double *__restrict__ bp = tm.bco + (kk - 1);
          const int top = min(c0 + SOS_CH - 1, NT);              // highest real level of the chunk
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
            // ---- from the ground upwards: levels top .. c0; level i uses layer i and S(i+1) ----
            int lv = top;
            const int full_hi = jdump ? c0 - 1 : ((top == NT) ? (NT & ~(EB - 1)) - 1 : top);   // levels c0 .. full_hi: whole blocks below NT
            for (; lv > full_hi; --lv) {                         // ragged head (at most EB levels; holds level NT)
              const int col = lv - c0;
              const double rw = raw(col);
              const double s = source(rw, col);
              if (lv == NT) z = bc;
              else z = z * attp[(size_t)lv * N] + (bp[(size_t)lv * N] * s + gp[(size_t)lv * N] * sedge);
              sedge = s; redge = rw;
              sAw[2 * acc_slot(e, col >> 1) + (col & 1)] = z;
              if (jrow) jrow[col] = s;
            }
            if (!u.lr) {
              // aerosol-only orders: c = pup*acc(i) + qup*acc(i+1), 3 FP64 instructions per level
              const double *__restrict__ pp = tm.pup + (kk - 1), *__restrict__ qp = tm.qup + (kk - 1);
              for (; lv >= c0; lv -= EB) {                       // whole blocks of EB levels, lv = highest level of the block
                const int cb = lv - (EB - 1) - c0;               // first column of the block
                double R[EB + 1], aa[EB], pw[EB], qw[EB];
#pragma unroll
                for (int j = 0; j < EB; ++j) {
                  const size_t l = (size_t)(c0 + cb + j) * N;
                  aa[j] = attp[l]; pw[j] = pp[l]; qw[j] = qp[l];
                }
#pragma unroll
                for (int p = 0; p < EB / 2; ++p) {
                  const double2 t = sA2[acc_slot(e, (cb >> 1) + p)];
                  R[2 * p] = t.x; R[2 * p + 1] = t.y;
                }
                R[EB] = redge;
#pragma unroll
                for (int p = EB / 2 - 1; p >= 0; --p) {          // the new field replaces the accumulators pair by pair
                  const double zo = z * aa[2 * p + 1] + (pw[2 * p + 1] * R[2 * p + 1] + qw[2 * p + 1] * R[2 * p + 2]);
                  z = zo * aa[2 * p] + (pw[2 * p] * R[2 * p] + qw[2 * p] * R[2 * p + 1]);
                  sAw2[acc_slot(e, (cb >> 1) + p)] = make_double2(z, zo);
                }
                redge = R[0];
              }
            } else
            for (; lv >= c0; lv -= EB) {
              const int cb = lv - (EB - 1) - c0;
              double S[EB + 1], aa[EB], gg[EB], bb[EB];
#pragma unroll
              for (int j = 0; j < EB; ++j) {
                const size_t l = (size_t)(c0 + cb + j) * N;
                aa[j] = attp[l]; gg[j] = gp[l]; bb[j] = bp[l];
              }
#pragma unroll
              for (int p = 0; p < EB / 2; ++p) {
                const double2 t = sA2[acc_slot(e, (cb >> 1) + p)];
                S[2 * p] = source(t.x, cb + 2 * p);
                S[2 * p + 1] = source(t.y, cb + 2 * p + 1);
              }
              S[EB] = sedge;
#pragma unroll
              for (int p = EB / 2 - 1; p >= 0; --p) {
                const double zo = z * aa[2 * p + 1] + (bb[2 * p + 1] * S[2 * p + 1] + gg[2 * p + 1] * S[2 * p + 2]);
                z = zo * aa[2 * p] + (bb[2 * p] * S[2 * p] + gg[2 * p] * S[2 * p + 1]);
                sAw2[acc_slot(e, (cb >> 1) + p)] = make_double2(z, zo);
              }
              sedge = S[0];
            }
          } else {
            // ---- from the top downwards: levels c0 .. top; level i uses layer i-1 and S(i-1) ----
            int lv = c0;
            const int n_full = jdump ? 0 : ((top + 1 - c0) / EB);  // whole blocks of EB levels in this chunk
            if (!u.lr) {
              const double *__restrict__ pp = tm.pdn + (kk - 1), *__restrict__ qp = tm.qdn + (kk - 1);
              for (int blk = 0; blk < n_full; ++blk, lv += EB) {
                const int cb = lv - c0;
                double R[EB + 1], aa[EB], pw[EB], qw[EB];
#pragma unroll
                for (int j = 0; j < EB; ++j) {
                  aa[j] = attp[(size_t)max(lv + j - 1, 0) * N];   // layer above level lv+j
                  const size_t l = (size_t)(lv + j) * N;
                  pw[j] = pp[l]; qw[j] = qp[l];                   // level-indexed (row 0 is zero: X(0) = 0)
                }
                R[0] = redge;
#pragma unroll
                for (int p = 0; p < EB / 2; ++p) {
                  const double2 t = sA2[acc_slot(e, (cb >> 1) + p)];
                  R[2 * p + 1] = t.x; R[2 * p + 2] = t.y;
                }
                if (lv == 0) aa[0] = 0.0;
#pragma unroll
                for (int p = 0; p < EB / 2; ++p) {
                  const double ze = z * aa[2 * p] + (pw[2 * p] * R[2 * p + 1] + qw[2 * p] * R[2 * p]);
                  z = ze * aa[2 * p + 1] + (pw[2 * p + 1] * R[2 * p + 2] + qw[2 * p + 1] * R[2 * p + 1]);
                  sAw2[acc_slot(e, (cb >> 1) + p)] = make_double2(ze, z);
                }
                redge = R[EB];
                sedge = sXd[cb + EB - 1] * R[EB];                 // the ragged tail of the profile continues with the generic form
              }
            } else
            for (int blk = 0; blk < n_full; ++blk, lv += EB) {
              const int cb = lv - c0;
              double S[EB + 1], aa[EB], gg[EB], bb[EB];
#pragma unroll
              for (int j = 0; j < EB; ++j) {
                const size_t l = (size_t)max(lv + j - 1, 0) * N;   // layer above level lv+j (level 0 has none: fixed below)
                aa[j] = attp[l]; gg[j] = gp[l]; bb[j] = bp[l];
              }
              S[0] = sedge;
#pragma unroll
              for (int p = 0; p < EB / 2; ++p) {
                const double2 t = sA2[acc_slot(e, (cb >> 1) + p)];
                S[2 * p + 1] = source(t.x, cb + 2 * p);
                S[2 * p + 2] = source(t.y, cb + 2 * p + 1);
              }
              if (lv == 0) { aa[0] = 0.0; gg[0] = 0.0; bb[0] = 0.0; }   // level 0: X = 0
#pragma unroll
              for (int p = 0; p < EB / 2; ++p) {
                const double ze = z * aa[2 * p] + (bb[2 * p] * S[2 * p + 1] + gg[2 * p] * S[2 * p]);
                z = ze * aa[2 * p + 1] + (bb[2 * p + 1] * S[2 * p + 2] + gg[2 * p + 1] * S[2 * p + 1]);
                sAw2[acc_slot(e, (cb >> 1) + p)] = make_double2(ze, z);
              }
              sedge = S[EB];
            }
            for (; lv <= top; ++lv) {                            // ragged tail
              const int col = lv - c0;
              const double rw = raw(col);
              const double s = source(rw, col);
              if (lv == 0) z = 0.0;
              else {
                const size_t l = (size_t)(lv - 1) * N;
                z = z * attp[l] + (bp[l] * s + gp[l] * sedge);
              }
              sedge = s; redge = rw;
              sAw[2 * acc_slot(e, col >> 1) + (col & 1)] = z;
              if (jrow) jrow[col] = s;
            }
          }
        }
        // ---- copy-out: the tile now holds X_n; whole rows go to HBM with one 512-byte store instruction per row
        //      (a row-per-thread store would touch 32 lines per instruction and stall the MMA warps' shared-memory loads) ----
        epi_bar();
        if (lane == 0) mbar_arrive(tab_empty);                   // the layer tables are free for the next chunk
        {
          const int ncol = min(SOS_CH, L - c0);                  // real levels of this chunk
          const int ew = e >> 5;
          for (int rl = ew; rl < u.R; rl += SW_EPI_WARPS) {
            if (u.g0 * 16 + rl >= 3 * N) break;                  // pad rows stay zero (rows ascend)
            const double2 t = sA2[acc_slot(rl, lane)];
            double *dst = xnext + SOS_XIDX(KP, u.r0 + rl, c0) + 2 * lane;
            if (2 * lane + 1 < ncol) *reinterpret_cast<double2 *>(dst) = t;
            else if (2 * lane < ncol) *dst = t.x;
          }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(acc_empty);
        ++accn;
      }
    }
  }
}

static size_t sweep_smem_bytes()
{
  return (size_t)SW_STG * SW_STAGE + (size_t)128 * SW_ACC_PITCH * 8 + (4 * SOS_CH + 2 * 72 + 3 * 80) * 8 + (2 * SW_STG + 24) * 8 +
         128;
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
  const size_t smem = sweep_smem_bytes();
  cudaMemsetAsync(work_counter, 0, sizeof(unsigned), st);
  const long long units = (long long)nitem * 2 * tiles_per_dir;
  const int grid = (int)std::min<long long>(num_sms, units);
  // 8 DMMA warps.  (A 16-warp instantiation, 16 rows x 32 levels per warp, was measured at 20.5 TFLOP/s against 28.0: the A
  // fragments are then loaded by twice as many warps and the register budget drops to 80.)
  static bool attr_set[64] = {};
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev >= 0 && dev < 64 && !attr_set[dev]) {                  // the opt-in to > 48 KB dynamic shared memory is per device
    if (cudaFuncSetAttribute(k_sweep<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return -1;
    attr_set[dev] = true;
  }
  k_sweep<8><<<grid, (8 + SW_EPI_WARPS + 1) * 32, smem, st>>>(items, terms, optics, ksets, list, count_ptr, nitem, tiles_per_dir, work_counter,
                                                             jdump);
  return 1;
}
