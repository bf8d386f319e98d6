// step_kernel.cu -- the fused hot kernel of libsosgpu.so (sm_100a).
//
// One launch advances every active (term, Fourier order) item by one scattering order:
//   J(level, mu) = XDEL(level) * (A_A X_{n-1})(level) + YDEL(level) * (A_R X_{n-1})(level)
//        = SOS_FSOURCE_ORDREIG (SOS_OS.F:2663-3017) as a dense FP64 contraction on DMMA tiles
//          (mma.sync.m8n8k4.f64 -> SASS DMMA.8x8x4).  Operands are staged into shared memory by the TMA
//          engine (cp.async.bulk + mbarrier complete_tx, SASS UBLKCP) through a 3-stage pipeline:
//          A_A is stored slab-major and pre-swizzled in HBM (k_pack), so one 16 KB bulk copy per k-slab
//          lands a conflict-free tile; the field chunk is copied row by row into padded rows.
//          The molecular part A_R has rank <= 4 per direction block (only the l=2 Legendre rows, SOS_OS.F:2859-2876):
//          it is carried as 4 extra functional rows V (T = V X, one extra DMMA per k-step and warp) and
//          expanded in the epilogue: J_R = u0 T0 + u(k) T_type.
//   X_n = SOS_INTEGR_EPOPT(J, boundary values)  (SOS_OS.F:2222-2357), run on the tile while it is still
//          on chip: the source function never goes to HBM;
//   boundary values at the ground for mu>0 (Lambert / BRDF-BPDF quadrature / flat Fresnel,
//          SOS_OS.F:1166-1239) are formed in the prologue from the previous order's downward field.
// ORDER1 variant: the source is the analytic first-order source (SOS_FSOURCE_ORDRE1, SOS_OS.F:2431-2565,
//          plus SOS_FSOURCE_DIFF_FRESNEL1 for a flat sea) and the boundary values are SOS_OS.F:970-992.
//
// CTA tile: up to 128 rows (8 warps x 16 rows) of one direction block x 64 levels; the level chunks are
// swept in the direction of propagation so the recurrence state stays in registers.  Two CTAs are
// resident per SM (<= 128 registers, ~83 KB shared memory each) so that one CTA's recurrence epilogue
// overlaps the other's DMMA main loop.
#include "sosgpu_internal.h"
#include <math.h>
#include <algorithm>
#include <cstdlib>
#include <cstdio>

#define STAGE_A_BYTES(rows) ((rows) * SOS_KB * 8)
#define STAGE_V_BYTES (8 * SOS_KB * 8)
#define STAGE_B_BYTES (SOS_KB * SOS_SB * 8)

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
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(ok)
        : "r"(addr), "r"(parity)
        : "memory");
  } while (!ok);
}
// TMA engine, non-tensor bulk copy global -> shared, completion counted on an mbarrier
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, unsigned bytes, unsigned long long *bar)
{
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory"); }

__device__ __forceinline__ void dmma8x8x4(double &c0, double &c1, double a, double b)
{
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(c0), "+d"(c1)
               : "d"(a), "d"(b));
}

// DMMAs of one k-slab for one warp: warp tile 16 rows x (8*NFR) levels (2 x NFR fragments of 8x8, K = 16 in 4 steps).
// NFR = column blocks (8 levels) of the chunk that hold real levels; k-steps >= ks_lim only touch zero padding.
template <int NFR, int LR>
__device__ __forceinline__ void slab_mma(double (&acc)[2][8][2], double (&tacc)[2][2], const double *__restrict__ a,
                                         const double *__restrict__ v, const double *__restrict__ b, bool own, int wr,
                                         int nwl, int gq, int tq, int ks_lim)
{
  const int swz = 4 * (gq & 3);
#pragma unroll
  for (int ks4 = 0; ks4 < SOS_KB / 4; ++ks4) {
    if (ks4 >= ks_lim) break;                                // uniform; false for 14 of 16 slabs
    const int kc = (ks4 * 4 + tq) ^ swz;
    if (own) {
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
    if (LR) {                                                // T = V X: this warp's share of the 8 column blocks
      const double av = v[kc];
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        const int ni = wr + u * nwl;
        if (ni < NFR) dmma8x8x4(tacc[u][0], tacc[u][1], av, b[ks4 * 4 * SOS_SB + ni * 8]);
      }
    }
  }
}

// EXPERIMENT (SOS_DBG & 1024): one "main loop" token per SM in global memory -- the two co-resident CTAs take turns in
// the DMMA main loop, so that one CTA's recurrence epilogue always runs under the other's main loop (anti-phase).
__device__ unsigned g_sm_token[256];

template <int LR, int ORDER1>
__global__ void __launch_bounds__(256, 2)
k_step(const ItemDev *__restrict__ items, const TermDev *__restrict__ terms, const OpticsDev *__restrict__ optics,
       const KsetDev *__restrict__ ksets, const int *__restrict__ list, int tiles_per_dir, int want_lr, int att_cap,
       double *__restrict__ jdump, int dbg)
{
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int nw_launch = blockDim.x >> 5;
  const int rows_max = nw_launch * 16;
  // ---- shared memory carve: [pipeline stages | aliased by the staging tile sJ] sT sG mbar ----
  const int stage_bytes = STAGE_A_BYTES(rows_max) + (LR ? STAGE_V_BYTES : 0) + STAGE_B_BYTES;
  const int pipe_bytes = ORDER1 ? 0 : SOS_STAGES * stage_bytes;
  const int sj_bytes = rows_max * SOS_SJ * 8;
  const int region = (pipe_bytes > sj_bytes ? pipe_bytes : sj_bytes);
  double *sJ = reinterpret_cast<double *>(smem_raw);
  double *sT = reinterpret_cast<double *>(smem_raw + region);           // [4][64] Rayleigh functionals of the chunk
  double *sG = sT + 4 * SOS_CH;                                         // [3N] ground values of the downward field
  unsigned long long *full = reinterpret_cast<unsigned long long *>(sG + 3 * 80);
  unsigned long long *empty = full + SOS_STAGES;
  unsigned long long *tabbar = empty + SOS_STAGES;                      // per-chunk layer tables landed
  double *sDt = reinterpret_cast<double *>(full + 2 * SOS_STAGES + 2);  // [<=66] layer optical thickness of the chunk
  double *sInv = sDt + 72;                                              // [<=66] reciprocal
  double *sXd = sInv + 72;                                              // [64] XDEL of the chunk's levels
  double *sYd = sXd + 72;                                               // [64] YDEL
  double *sU = sYd + 72;                                                // [128] urow of the tile's rows (LR)
  double *sCh = sU + 128;                                               // [64] CH (and [64] CF) of the chunk (ORDER1)
  double *sC = sCh + 2 * 72;                                            // [4][128] c1, c2, fz1, fz2 of the tile's rows (ORDER1)
  double *sAtt = sC + 4 * 128;                                          // [<=66][N] exp(-dtau/mu_k) (when att_cap > 0)

  const int per_item = 2 * tiles_per_dir;
  const int ii = blockIdx.x / per_item, t = blockIdx.x % per_item;
  const int item = list ? list[ii] : ii;
  const ItemDev it = items[item];
  const KsetDev ks = ksets[it.kset];
  if (!ORDER1 && (ks.dual != want_lr)) return;                 // handled by the other instantiation
  const TermDev tm = terms[it.term];
  const OpticsDev &op = optics[tm.optics];
  const int N = op.nbmu, HB = op.HB, KP = op.KP, NT = tm.nt, L = NT + 1, LP = tm.LP;
  const int dir = t / tiles_per_dir, tile = t % tiles_per_dir;
  const int groups = HB >> 4;
  const int gpt = (groups + tiles_per_dir - 1) / tiles_per_dir;
  const int g0 = tile * gpt;
  if (g0 >= groups) return;
  const int ng = min(gpt, groups - g0);                          // 16-row groups in this tile
  const int R = ng * 16;
  const int r0 = dir * HB + g0 * 16;                             // first packed row of the tile
  const int tid = threadIdx.x, lane = tid & 31, wr = tid >> 5;
  const int gq = lane >> 2, tq = lane & 3;
  const bool up = (dir == 0);

  const double *__restrict__ xprev = it.x[it.n & 1];
  double *__restrict__ xnext = ORDER1 ? it.x[1] : it.x[(it.n + 1) & 1];

  if (tid == 0) {
    for (int s = 0; s < SOS_STAGES; ++s) { mbar_init(full + s, 1); mbar_init(full + SOS_STAGES + s, nw_launch); }
    mbar_init(tabbar, 1);
    fence_proxy_async();                                         // make the inits visible to the async proxy
  }

  if (LR && tid < R) sU[tid] = (r0 + tid - dir * HB < 3 * N) ? ks.urow[r0 + tid] : 0.0;
  if (ORDER1 && tid < R) {
    sC[tid] = ks.c1[r0 + tid]; sC[128 + tid] = ks.c2[r0 + tid]; sC[256 + tid] = ks.fz1[r0 + tid]; sC[384 + tid] = ks.fz2[r0 + tid];
  }
  // ---------------- prologue: boundary value of this thread's row (mu > 0 rows only) ----------------
  const int myrow = r0 + tid;                                    // scan threads: tid < R
  const int q = (tid < R) ? (myrow - dir * HB) : 3 * N;
  const bool rowvalid = (tid < R) && (q < 3 * N);
  const int so = rowvalid ? q / N : 0, kk = rowvalid ? q % N + 1 : 1;
  const double mu = op.rmu[kk + N];
  double bc = 0.0;
  if (up) {
    if (ORDER1) {
      if (rowvalid) {                                            // SOS_OS.F:970-992
        if (so == 0 && !(op.ro == 0.0 || it.is != 0)) bc = -op.ro * op.tab * tm.eground;
        if (op.imat_surf == 1) {
          const float *rs = op.surf + (size_t)it.is * 9 * N * N;
          const double rr = tm.eground / mu;
          double rv = (double)rs[(size_t)(so * 3) * N * N + (size_t)(kk - 1) * N + (op.n0 - 1)];
          if (op.ipolar == 0 && so != 0) rv = 0.0;
          bc = (so == 0) ? bc + rv * rr : rv * rr;
        }
      }
    } else {
      // ground values of the previous order's downward field: G[st*N + j-1] = X_{n-1}(NT, -j)
      for (int c = tid; c < 3 * N; c += blockDim.x) sG[c] = xprev[SOS_XIDX(KP, HB + c, NT)];
      __syncthreads();
      if (rowvalid) {
        double xr = 0.0;
        if (!(op.ro == 0.0 || it.is != 0)) {                     // Lambert (SOS_OS.F:1177-1190)
          double lsol = 0.0;
          for (int j = 1; j <= N; ++j) lsol = lsol + op.ga[j + N] * sG[j - 1] * op.rmu[j + N];
          lsol = 2 * lsol * op.ro;
          xr = lsol;
          if (so == 0) bc = lsol;
        }
        if (op.imat_surf == 1) {                                 // BRDF / BPDF quadrature (SOS_OS.F:1194-1220)
          const float *rs = op.surf + (size_t)it.is * 9 * N * N + (size_t)(so * 3) * N * N + (size_t)(kk - 1) * N;
          const bool pol = (op.ipolar != 0);
          double acc = 0.0;
          for (int j = 1; j <= N; ++j) {
            double r1 = (double)rs[j - 1], r2 = (double)rs[(size_t)N * N + j - 1], r3 = (double)rs[(size_t)2 * N * N + j - 1];
            if (!pol) { if (so == 0) { r2 = 0.0; r3 = 0.0; } else { r1 = 0.0; r2 = 0.0; r3 = 0.0; } }
            acc = acc + op.ga[j + N] * (sG[j - 1] * r1 + sG[N + j - 1] * r2 + sG[2 * N + j - 1] * r3);
          }
          const double rrmu = 2 / mu;
          bc = (so == 0) ? acc * rrmu + xr : acc * rrmu;
        }
        if (op.ifresnel == 1) {                                  // flat sea (SOS_OS.F:1225-1239)
          const double gi = sG[kk - 1], gqv = sG[N + kk - 1], gu = sG[2 * N + kk - 1];
          if (so == 0) bc = bc + op.f11[kk] * gi + op.f12[kk] * gqv;
          else if (so == 1) bc = bc + op.f12[kk] * gi + op.f11[kk] * gqv;
          else bc = bc + op.f33[kk] * gu;
        }
      }
    }
  }
  __syncthreads();                                               // mbarrier inits visible to all threads

  // ---------------- level chunks, swept in the direction of propagation ----------------
  const int n_chunk = (L + SOS_CH - 1) / SOS_CH;
  const int n_slab = KP / SOS_KB;
  // slab-major pre-swizzled operands (k_pack): A[(slab*KP + row)*16 + ...], V[dir][(slab*8 + r)*16 + ...]
  const double *__restrict__ Ag = ks.apackA + (size_t)r0 * SOS_KB;
  const double *__restrict__ Vg = LR ? ks.vpack + (size_t)dir * 8 * KP : nullptr;

  double acc[2][8][2];
  double tacc[2][2];                                             // T = V X blocks of this warp (ni = wr + u*nw_launch, nw_launch >= 4)
  double z = 0.0;                                                 // recurrence state of this thread's row (pass 2)
  double scarry = 0.0;                                            // source function at the neighbouring level of the previous chunk (pass 1)
  unsigned it_count = 0;                                          // global slab counter (mbarrier phases)

#ifdef SOS_PHASE_TIMING   // build with -DSOS_PHASE_TIMING and run with SOS_DBG=16: cycles per phase as seen by thread 0
  long long tph[6] = {0, 0, 0, 0, 0, 0}, tmark = clock64();
  const long long tstart = tmark;
#define TPH(i) do { if (dbg & 16) { const long long now_ = clock64(); tph[i] += now_ - tmark; tmark = now_; } } while (0)
#else
#define TPH(i) do { } while (0)
#endif
  for (int chunk = 0; chunk < n_chunk; ++chunk) {
    const int ci = up ? (n_chunk - 1 - chunk) : chunk;
    const int c0 = ci * SOS_CH;
    // layers touched by this chunk's recurrence: [c0-1, c0+63] clipped; even start so the bulk copies are 16-byte aligned
    const int lb_al = max(c0 - 1, 0) & ~1;
    const bool att_staged = (att_cap >= 66 * N);
    if (tid == 0) {
      const int le = min(c0 + SOS_CH - 1, NT - 1);
      const int nrow = (le - lb_al + 2) & ~1;                      // even count (tables are padded by 2 rows)
      fence_proxy_async();
      const int nlev = (min(SOS_CH, L - c0) + 1) & ~1;             // levels of the chunk, even count
      mbar_expect_tx(tabbar, (unsigned)(nrow * 16 + nlev * 16 + (ORDER1 ? nlev * 16 : 0) + (att_staged ? nrow * N * 8 : 0)));
      if (ORDER1) {
        bulk_g2s(sCh, tm.ch + c0, (unsigned)(nlev * 8), tabbar);
        bulk_g2s(sCh + 72, tm.cf + c0, (unsigned)(nlev * 8), tabbar);
      }
      bulk_g2s(sXd, tm.xdel + c0, (unsigned)(nlev * 8), tabbar);
      bulk_g2s(sYd, tm.ydel + c0, (unsigned)(nlev * 8), tabbar);
      bulk_g2s(sDt, tm.dt + lb_al, (unsigned)(nrow * 8), tabbar);
      bulk_g2s(sInv, tm.inv + lb_al, (unsigned)(nrow * 8), tabbar);
      if (att_staged) bulk_g2s(sAtt, tm.att + (size_t)lb_al * N, (unsigned)(nrow * N * 8), tabbar);
    }

    if (!ORDER1) {
      const unsigned tx = (unsigned)(R * SOS_KB * 8 + (LR ? STAGE_V_BYTES : 0) + STAGE_B_BYTES);
      auto issue = [&](int slab, unsigned cnt) {                   // executed by thread 0 only
        const int stage = cnt % SOS_STAGES;
        unsigned char *sp = smem_raw + stage * stage_bytes;
        unsigned long long *bar = full + stage;
        mbar_expect_tx(bar, tx);
        bulk_g2s(sp, Ag + (size_t)slab * KP * SOS_KB, (unsigned)(R * SOS_KB * 8), bar);
        sp += STAGE_A_BYTES(rows_max);
        if (LR) { bulk_g2s(sp, Vg + (size_t)slab * 8 * SOS_KB, STAGE_V_BYTES, bar); sp += STAGE_V_BYTES; }
        // chunk-major padded field layout: the 16 x 64 slab (pitch SOS_SB) is one contiguous block
        bulk_g2s(sp, xprev + SOS_XIDX(KP, slab * SOS_KB, c0), STAGE_B_BYTES, bar);
      };
      if (tid == 0) {
        for (int s = 0; s < SOS_STAGES - 1 && s < n_slab; ++s) issue(s, it_count + s);
      }
      if (dbg & 1024) {
        if (tid == 0) {
          unsigned smid;
          asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
          while (atomicCAS(&g_sm_token[smid], 0u, 1u) != 0u) __nanosleep(64);
        }
        __syncthreads();
      }
      TPH(5);
#pragma unroll
      for (int mi = 0; mi < 2; ++mi)
#pragma unroll
        for (int ni = 0; ni < 8; ++ni) { acc[mi][ni][0] = 0.0; acc[mi][ni][1] = 0.0; }
      tacc[0][0] = tacc[0][1] = tacc[1][0] = tacc[1][1] = 0.0;

      const int nfr = (min(SOS_CH, L - c0) + 7) >> 3;              // column blocks (8 levels) that hold real levels
      for (int slab = 0; slab < n_slab; ++slab) {
        const unsigned cnt = it_count + slab;
        const int stage = cnt % SOS_STAGES;
        // refill the stage used one iteration ago once every warp has released it (empty barrier), no CTA barrier
        if (tid == 0 && slab + SOS_STAGES - 1 < n_slab) {
          if (slab >= 1) mbar_wait(empty + (cnt - 1) % SOS_STAGES, ((cnt - 1) / SOS_STAGES) & 1);
          issue(slab + SOS_STAGES - 1, cnt + SOS_STAGES - 1);
        }
        mbar_wait(full + stage, (cnt / SOS_STAGES) & 1);
        if (wr < ng || LR) {
          const unsigned char *sp = smem_raw + stage * stage_bytes;
          const double *a = reinterpret_cast<const double *>(sp) + (wr * 16 + gq) * SOS_KB;
          const double *v = reinterpret_cast<const double *>(sp + STAGE_A_BYTES(rows_max)) + gq * SOS_KB;
          const double *b = reinterpret_cast<const double *>(sp + STAGE_A_BYTES(rows_max) + (LR ? STAGE_V_BYTES : 0)) + tq * SOS_SB + gq;
          // k-steps that only touch the zero padding of a direction block contribute nothing: ks_lim of this slab
          const int kq = (slab * SOS_KB) % HB;
          const int ks_lim = (kq + SOS_KB <= 3 * N) ? 4 : max(0, (3 * N - kq + 3) >> 2);
          const bool own = wr < ng;
          // column blocks (8 levels) beyond the profile are skipped: one unrolled, predicate-free body per count
          switch (nfr) {
            case 8: slab_mma<8, LR>(acc, tacc, a, v, b, own, wr, nw_launch, gq, tq, ks_lim); break;
            case 7: slab_mma<7, LR>(acc, tacc, a, v, b, own, wr, nw_launch, gq, tq, ks_lim); break;
            case 6: slab_mma<6, LR>(acc, tacc, a, v, b, own, wr, nw_launch, gq, tq, ks_lim); break;
            case 5: slab_mma<5, LR>(acc, tacc, a, v, b, own, wr, nw_launch, gq, tq, ks_lim); break;
            case 4: slab_mma<4, LR>(acc, tacc, a, v, b, own, wr, nw_launch, gq, tq, ks_lim); break;
            case 3: slab_mma<3, LR>(acc, tacc, a, v, b, own, wr, nw_launch, gq, tq, ks_lim); break;
            case 2: slab_mma<2, LR>(acc, tacc, a, v, b, own, wr, nw_launch, gq, tq, ks_lim); break;
            default: slab_mma<1, LR>(acc, tacc, a, v, b, own, wr, nw_launch, gq, tq, ks_lim); break;
          }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(empty + stage);                 // this warp no longer reads the stage
      }
      it_count += n_slab;
      __syncthreads();                                             // pipeline drained: stages may be reused as sJ
      if ((dbg & 1024) && tid == 0) {
        unsigned smid;
        asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
        atomicExch(&g_sm_token[smid], 0u);
      }
      TPH(0);
      mbar_wait(tabbar, chunk & 1);                                // XDEL/YDEL and layer tables of this chunk have landed
      if (LR && gq < 4) {
#pragma unroll
        for (int u = 0; u < 2; ++u) {
          const int ni = wr + u * nw_launch;
          if (ni < 8) { sT[gq * SOS_CH + ni * 8 + 2 * tq] = tacc[u][0]; sT[gq * SOS_CH + ni * 8 + 2 * tq + 1] = tacc[u][1]; }
        }
      }
      if (LR) __syncthreads();
    }

    // ---------------- tile epilogue: source function -> staging tile ----------------
    if (ORDER1) {
      __syncthreads();                                           // previous chunk's write-out has finished
      mbar_wait(tabbar, chunk & 1);
      {
        // one column per thread, rows strided by blockDim/64: level-dependent factors are loaded once
        const int col = tid & 63, level = c0 + col;
        const bool lv_ok = level < L;
        const double chv = lv_ok ? sCh[col] : 0.0, xdv = lv_ok ? sXd[col] : 0.0, ydv = lv_ok ? sYd[col] : 0.0;
        const bool fr_on = (op.ifresnel == 1) && lv_ok && (up ? (level <= NT - 1) : (level >= 1));
        const double cfv = fr_on ? sCh[72 + col] : 0.0;
        for (int rowl = tid >> 6; rowl < R; rowl += (blockDim.x >> 6)) {
          // SOS_FSOURCE_ORDRE1 (SOS_OS.F:2553-2560): ATTDIR*(S2*PCAER + S1*PCRAY)
          double v = __dmul_rn(chv, __dadd_rn(__dmul_rn(sC[128 + rowl], xdv), __dmul_rn(sC[rowl], ydv)));
          // SOS_FSOURCE_DIFF_FRESNEL1 (SOS_OS.F:3224-3292)
          if (fr_on) v = v + cfv * (sC[384 + rowl] * xdv + sC[256 + rowl] * ydv);
          sJ[rowl * SOS_SJ + col] = lv_ok ? v : 0.0;
        }
      }
    } else {
      if (wr < ng) {
        // Rayleigh expansion coefficients of this thread's two rows (SOS_OS.F:2859-2876 in factored form)
        double u0[2] = {0.0, 0.0}, us[2] = {0.0, 0.0};
        int ty[2] = {0, 0};
        if (LR) {
#pragma unroll
          for (int mi = 0; mi < 2; ++mi) {
            const int qr = g0 * 16 + wr * 16 + mi * 8 + gq;        // row within the direction block
            if (qr < 3 * N) {
              ty[mi] = qr / N;
              us[mi] = sU[wr * 16 + mi * 8 + gq];
              u0[mi] = (ty[mi] == 0) ? 1.0 : 0.0;
            }
          }
        }
#pragma unroll
        for (int ni = 0; ni < 8; ++ni) {
#pragma unroll
          for (int e = 0; e < 2; ++e) {
            const int col = ni * 8 + 2 * tq + e;
            const int level = c0 + col;
            double xd = 0.0, yd = 0.0;
            if (level < L) { xd = sXd[col]; if (LR) yd = sYd[col]; }
#pragma unroll
            for (int mi = 0; mi < 2; ++mi) {
              double v = (level < L) ? xd * acc[mi][ni][e] : 0.0;
              if (LR && level < L) {
                const double tsel = sT[(ty[mi] + 1) * SOS_CH + col];
                v = v + yd * (u0[mi] * sT[col] + us[mi] * tsel);
              }
              sJ[(wr * 16 + mi * 8 + gq) * SOS_SJ + col] = v;
            }
          }
        }
      }
    }
    __syncthreads();

    if (jdump) {                                                 // test hook: expose the source function
      for (int idx = tid; idx < R * SOS_CH; idx += blockDim.x) {
        const int rowl = idx >> 6, col = idx & 63;
        const int level = c0 + col;
        if (level < L) jdump[(size_t)(r0 + rowl) * LP + level] = sJ[rowl * SOS_SJ + col];
      }
      __syncthreads();
    }

    TPH(1);
    // ---------------- SOS_INTEGR_EPOPT on the tile (SOS_OS.F:2279-2310 up, 2320-2354 down) ----------------
    // The layer update is z <- z*a + c with c = (1-a)*(A*mu + S) -/+ A*(a*dtau), A = dS/dtau independent of z.
    // Pass 1 (all threads, two per row, 32 levels each): c replaces S in the staging tile.
    // Pass 2 (one thread per row): the bare recurrence, one dependent FMA per level.
    {
      const int row2 = tid >> 1, half = tid & 1;
      const int q2 = (row2 < R) ? (r0 + row2 - dir * HB) : 3 * N;
      const int hi = min(c0 + SOS_CH - 1, NT);
      const int ncolv = hi - c0 + 1;                             // columns of this chunk that hold real levels
      const bool valid2 = q2 < 3 * N;
      double *js = sJ + (valid2 ? row2 : 0) * SOS_SJ;
      const int k2 = valid2 ? q2 % N + 1 : 1;
      const double mu2 = op.rmu[k2 + N];
      const double *att = att_staged ? (sAtt + (k2 - 1) - (size_t)lb_al * N) : (tm.att + (k2 - 1));
      const double *dtp = sDt - lb_al, *ivp = sInv - lb_al;
      mbar_wait(tabbar, chunk & 1);                              // layer tables of this chunk have landed
      const int lo_c = 32 * half, hi_c = min(32 * half + 31, ncolv - 1);
      // boundary values are read before anybody overwrites them: column i needs S(i) and S(i+1) (up) / S(i-1) (down)
      double bnd = scarry, s_edge = 0.0;
      if (valid2) {
        if (up) {
          if (half == 0) bnd = (ncolv > 32) ? js[32] : 0.0;
          s_edge = (hi_c >= lo_c) ? js[lo_c] : 0.0;
        } else {
          if (half == 1) bnd = js[31];
          s_edge = (hi_c >= lo_c) ? js[hi_c] : 0.0;
        }
      }
      __syncwarp();
      if (valid2) {
        // four independent columns per iteration (the loop is latency bound: few warps, long FP64 chains)
        if (up) {
          for (int i = lo_c; i <= hi_c; i += 4) {                  // ascending: S(i+1) is still unmodified
            double sv[5], cv[4];
#pragma unroll
            for (int u = 0; u < 5; ++u) { const int c = i + u; sv[u] = (c <= hi_c) ? js[c] : ((c == hi_c + 1) ? bnd : 0.0); }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
              const int lv = min(c0 + i + u, NT - 1);
              const double a = att[(size_t)lv * N], dl = dtp[lv], iv = ivp[lv];
              const double A = (sv[u + 1] - sv[u]) * iv;
              cv[u] = (1.0 - a) * (A * mu2 + sv[u]) - A * (a * dl);
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) if (i + u <= hi_c && c0 + i + u < NT) js[i + u] = cv[u];
          }
        } else {
          const double rmuk = -mu2;
          for (int i = hi_c; i >= lo_c; i -= 4) {                  // descending: S(i-1) is still unmodified
            double sv[5], cv[4];
#pragma unroll
            for (int u = 0; u < 5; ++u) { const int c = i - u; sv[u] = (c >= lo_c) ? js[c] : ((c == lo_c - 1) ? bnd : 0.0); }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
              const int lv = max(c0 + i - u, 1);
              const double a = att[(size_t)(lv - 1) * N], dl = dtp[lv - 1], iv = ivp[lv - 1];
              const double A = (sv[u] - sv[u + 1]) * iv;
              cv[u] = (1.0 - a) * (A * rmuk + sv[u]) + A * (a * dl);
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) if (i - u >= lo_c && c0 + i - u > 0) js[i - u] = cv[u];
          }
        }
      }
      // the first (up) / last (down) column's S is the neighbour the next chunk needs: hand it to the partner thread
      const double give = __shfl_sync(0xffffffffu, s_edge, up ? (lane & ~1) : (lane | 1));
      if (half == (up ? 1 : 0)) scarry = give;
    }
    __syncthreads();
    TPH(2);
    if (rowvalid) {
      double *js = sJ + tid * SOS_SJ;
      const double *att = att_staged ? (sAtt + (kk - 1) - (size_t)lb_al * N) : (tm.att + (kk - 1));
      const int hi = min(c0 + SOS_CH - 1, NT);
      if (up) {
        int level = hi;
        if (level == NT) { z = bc; js[level - c0] = z; --level; }
        double aq[4];
#pragma unroll
        for (int p = 0; p < 4; ++p) aq[p] = (level - p >= c0) ? att[(size_t)(level - p) * N] : 0.0;
        for (; level >= c0; level -= 4) {
#pragma unroll
          for (int p = 0; p < 4; ++p) {
            const int lv = level - p;
            const double ac = aq[p];
            if (lv - 4 >= c0) aq[p] = att[(size_t)(lv - 4) * N];
            if (lv >= c0) { z = z * ac + js[lv - c0]; js[lv - c0] = z; }
          }
        }
      } else {
        int level = c0;
        if (level == 0) { z = 0.0; js[0] = 0.0; ++level; }
        double aq[4];
#pragma unroll
        for (int p = 0; p < 4; ++p) aq[p] = (level + p <= hi) ? att[(size_t)(level + p - 1) * N] : 0.0;
        for (; level <= hi; level += 4) {
#pragma unroll
          for (int p = 0; p < 4; ++p) {
            const int lv = level + p;
            const double ac = aq[p];
            if (lv + 4 <= hi) aq[p] = att[(size_t)(lv + 3) * N];
            if (lv <= hi) { z = z * ac + js[lv - c0]; js[lv - c0] = z; }
          }
        }
      }
    }
    __syncthreads();

    TPH(3);
    // ---------------- write the new field tile (coalesced along levels) ----------------
    for (int idx = tid; idx < R * SOS_CH; idx += blockDim.x) {
      const int rowl = idx >> 6, col = idx & 63;
      const int level = c0 + col, row = r0 + rowl;
      if (level < L && (row - dir * HB) < 3 * N) xnext[SOS_XIDX(KP, row, level)] = sJ[rowl * SOS_SJ + col];
    }
    __syncthreads();                                             // sJ is free again (next chunk's TMA may overwrite it)
    TPH(4);
  }
#ifdef SOS_PHASE_TIMING
  if ((dbg & 16) && tid == 0 && (blockIdx.x % 997) == 0)
    printf("cta %d chunks %d slabs %d R %d: mainloop %lld store %lld pass1 %lld pass2 %lld writeout %lld token+restart %lld total %lld\n", (int)blockIdx.x, n_chunk,
           n_slab, R, tph[0], tph[1], tph[2], tph[3], tph[4], tph[5], clock64() - tstart);
#endif
#undef TPH
}

// =================================================================================================
// k_step2 -- EXPERIMENTAL (SOS_DBG & 128), parity-green but slower than k_step as of round 1 (16.4 vs 22.6 TFLOP/s):
// same operator as k_step<LR, 0>, different schedule:
//   * the TMA pipeline runs over the linear sequence (chunk, k-slab) and never drains between level chunks;
//   * the recurrence epilogue works on the DMMA accumulators IN REGISTERS: source function, layer constants c and the
//     serial recurrence z <- z*a + c (handed from lane to lane of a quad by shuffles, in the reference's level order)
//     never go through a shared staging tile, and the new field is stored straight from the accumulator layout
//     (16-byte stores, a quad writes 64 contiguous bytes of a row);
//   * no CTA barrier per chunk (LR: one, for the 4 x 64 Rayleigh functionals): warps drift apart, so one warp's
//     epilogue runs under the other warps' DMMAs of the next chunk.
// Per-chunk layer tables are single-buffered behind a full/free mbarrier pair.
// STG: pipeline stages; ATTS: attenuation table of the chunk staged in shared memory (else read from global memory,
// whose pad rows >= NT are zero)
template <int LR, int STG, int ATTS>
__global__ void __launch_bounds__(256, STG > 4 ? 1 : 2)
k_step2(const ItemDev *__restrict__ items, const TermDev *__restrict__ terms, const OpticsDev *__restrict__ optics,
        const KsetDev *__restrict__ ksets, const int *__restrict__ list, int tiles_per_dir, int want_lr, int att_cap)
{
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int nthr = blockDim.x;
  const int nw_launch = nthr >> 5;
  const int rows_max = nw_launch * 16;
  const int stage_bytes = STAGE_A_BYTES(rows_max) + (LR ? STAGE_V_BYTES : 0) + STAGE_B_BYTES;
  double *sT = reinterpret_cast<double *>(smem_raw + STG * stage_bytes);   // [2][4][64] Rayleigh functionals (double-buffered)
  double *sG = sT + 2 * 4 * SOS_CH;                                     // [3N] ground values of the downward field
  unsigned long long *full = reinterpret_cast<unsigned long long *>(sG + 3 * 80);
  unsigned long long *empty = full + STG;
  unsigned long long *tabfull = empty + STG;                     // layer tables of the chunk have landed
  unsigned long long *tabfree = tabfull + 1;                            // every warp is done with the chunk's tables
  double *sDt = reinterpret_cast<double *>(full + 2 * STG + 2);  // [<=66] layer optical thickness of the chunk
  double *sInv = sDt + 72;
  double *sXd = sInv + 72;
  double *sYd = sXd + 72;
  double *sU = sYd + 72;                                                // [128] urow of the tile's rows (LR)
  double *sBc = sU + 128;                                               // [128] ground boundary value of the tile's rows
  double *sAtt = sBc + 128;                                             // [<=66][N] exp(-dtau/mu_k) (when att_cap > 0)

  const int per_item = 2 * tiles_per_dir;
  const int ii = blockIdx.x / per_item, t = blockIdx.x % per_item;
  const int item = list ? list[ii] : ii;
  const ItemDev it = items[item];
  const KsetDev ks = ksets[it.kset];
  if (ks.dual != want_lr) return;                                // handled by the other instantiation
  const TermDev tm = terms[it.term];
  const OpticsDev &op = optics[tm.optics];
  const int N = op.nbmu, HB = op.HB, KP = op.KP, NT = tm.nt, L = NT + 1;
  const int dir = t / tiles_per_dir, tile = t % tiles_per_dir;
  const int groups = HB >> 4;
  const int gpt = (groups + tiles_per_dir - 1) / tiles_per_dir;
  const int g0 = tile * gpt;
  if (g0 >= groups) return;
  const int ng = min(gpt, groups - g0);
  const int R = ng * 16;
  const int r0 = dir * HB + g0 * 16;
  const int tid = threadIdx.x, lane = tid & 31, wr = tid >> 5;
  const int gq = lane >> 2, tq = lane & 3;
  const bool up = (dir == 0);
  const double *__restrict__ xprev = it.x[it.n & 1];
  double *__restrict__ xnext = it.x[(it.n + 1) & 1];

  if (tid == 0) {
    for (int s = 0; s < STG; ++s) { mbar_init(full + s, 1); mbar_init(empty + s, nw_launch); }
    mbar_init(tabfull, 1);
    mbar_init(tabfree, nw_launch);
    fence_proxy_async();
  }
  if (LR && tid < R) sU[tid] = (r0 + tid - dir * HB < 3 * N) ? ks.urow[r0 + tid] : 0.0;

  // ---------------- prologue: ground boundary value of row tid (mu > 0 rows only), as in k_step ----------------
  {
    const int q = (tid < R) ? (r0 + tid - dir * HB) : 3 * N;
    const bool rowvalid = (tid < R) && (q < 3 * N);
    const int so = rowvalid ? q / N : 0, kk = rowvalid ? q % N + 1 : 1;
    const double mu = op.rmu[kk + N];
    double bc = 0.0;
    if (up) {
      for (int c = tid; c < 3 * N; c += nthr) sG[c] = xprev[SOS_XIDX(KP, HB + c, NT)];
      __syncthreads();
      if (rowvalid) {
        double xr = 0.0;
        if (!(op.ro == 0.0 || it.is != 0)) {                     // Lambert (SOS_OS.F:1177-1190)
          double lsol = 0.0;
          for (int j = 1; j <= N; ++j) lsol = lsol + op.ga[j + N] * sG[j - 1] * op.rmu[j + N];
          lsol = 2 * lsol * op.ro;
          xr = lsol;
          if (so == 0) bc = lsol;
        }
        if (op.imat_surf == 1) {                                 // BRDF / BPDF quadrature (SOS_OS.F:1194-1220)
          const float *rs = op.surf + (size_t)it.is * 9 * N * N + (size_t)(so * 3) * N * N + (size_t)(kk - 1) * N;
          const bool pol = (op.ipolar != 0);
          double acc = 0.0;
          for (int j = 1; j <= N; ++j) {
            double r1 = (double)rs[j - 1], r2 = (double)rs[(size_t)N * N + j - 1], r3 = (double)rs[(size_t)2 * N * N + j - 1];
            if (!pol) { if (so == 0) { r2 = 0.0; r3 = 0.0; } else { r1 = 0.0; r2 = 0.0; r3 = 0.0; } }
            acc = acc + op.ga[j + N] * (sG[j - 1] * r1 + sG[N + j - 1] * r2 + sG[2 * N + j - 1] * r3);
          }
          const double rrmu = 2 / mu;
          bc = (so == 0) ? acc * rrmu + xr : acc * rrmu;
        }
        if (op.ifresnel == 1) {                                  // flat sea (SOS_OS.F:1225-1239)
          const double gi = sG[kk - 1], gqv = sG[N + kk - 1], gu = sG[2 * N + kk - 1];
          if (so == 0) bc = bc + op.f11[kk] * gi + op.f12[kk] * gqv;
          else if (so == 1) bc = bc + op.f12[kk] * gi + op.f11[kk] * gqv;
          else bc = bc + op.f33[kk] * gu;
        }
      }
    }
    if (tid < 128) sBc[tid] = bc;
  }
  __syncthreads();                                               // mbarrier inits, sU, sBc visible

  const int n_chunk = (L + SOS_CH - 1) / SOS_CH;
  const int n_slab = KP / SOS_KB;
  const int n_iter = n_chunk * n_slab;                           // linear (chunk, slab) sequence of the pipeline
  const double *__restrict__ Ag = ks.apackA + (size_t)r0 * SOS_KB;
  const double *__restrict__ Vg = LR ? ks.vpack + (size_t)dir * 8 * KP : nullptr;
  const bool att_staged = ATTS != 0;
  const unsigned tx = (unsigned)(R * SOS_KB * 8 + (LR ? STAGE_V_BYTES : 0) + STAGE_B_BYTES);

  auto issue = [&](int qi) {                                     // thread 0 only: loads of linear iteration qi
    const int ch = qi / n_slab, slab = qi - ch * n_slab;
    const int c0i = (up ? (n_chunk - 1 - ch) : ch) * SOS_CH;
    unsigned char *sp = smem_raw + (qi % STG) * stage_bytes;
    unsigned long long *bar = full + (qi % STG);
    mbar_expect_tx(bar, tx);
    bulk_g2s(sp, Ag + (size_t)slab * KP * SOS_KB, (unsigned)(R * SOS_KB * 8), bar);
    sp += STAGE_A_BYTES(rows_max);
    if (LR) { bulk_g2s(sp, Vg + (size_t)slab * 8 * SOS_KB, STAGE_V_BYTES, bar); sp += STAGE_V_BYTES; }
    bulk_g2s(sp, xprev + SOS_XIDX(KP, slab * SOS_KB, c0i), STAGE_B_BYTES, bar);
  };
  auto issue_tables = [&](int ch) {                              // thread 0 only: layer tables of chunk ch
    const int c0i = (up ? (n_chunk - 1 - ch) : ch) * SOS_CH;
    const int lb = max(c0i - 1, 0) & ~1;
    const int le = min(c0i + SOS_CH - 1, NT - 1);
    const int nrow = (le - lb + 2) & ~1;
    const int nlev = (min(SOS_CH, L - c0i) + 1) & ~1;
    fence_proxy_async();
    mbar_expect_tx(tabfull, (unsigned)(nrow * 16 + nlev * 16 + (att_staged ? nrow * N * 8 : 0)));
    bulk_g2s(sXd, tm.xdel + c0i, (unsigned)(nlev * 8), tabfull);
    bulk_g2s(sYd, tm.ydel + c0i, (unsigned)(nlev * 8), tabfull);
    bulk_g2s(sDt, tm.dt + lb, (unsigned)(nrow * 8), tabfull);
    bulk_g2s(sInv, tm.inv + lb, (unsigned)(nrow * 8), tabfull);
    if (att_staged) bulk_g2s(sAtt, tm.att + (size_t)lb * N, (unsigned)(nrow * N * 8), tabfull);
  };
  if (tid == 0) {
    issue_tables(0);
    for (int s = 0; s < STG - 1 && s < n_iter; ++s) issue(s);
  }

  // rows of this thread in the accumulator layout: local row wr*16 + mi*8 + gq
  const bool own = wr < ng;
  bool rvalid[2];
  double mu_r[2], bc_r[2], u0[2] = {0.0, 0.0}, us[2] = {0.0, 0.0};
  int kidx[2], ty[2] = {0, 0};
#pragma unroll
  for (int mi = 0; mi < 2; ++mi) {
    const int rl = wr * 16 + mi * 8 + gq;
    const int qr = g0 * 16 + rl;                                 // row within the direction block
    rvalid[mi] = own && (qr < 3 * N);
    kidx[mi] = rvalid[mi] ? qr % N + 1 : 1;
    mu_r[mi] = op.rmu[kidx[mi] + N];
    bc_r[mi] = own ? sBc[rl & 127] : 0.0;
    if (LR && rvalid[mi]) { ty[mi] = qr / N; us[mi] = sU[rl]; u0[mi] = (ty[mi] == 0) ? 1.0 : 0.0; }
  }
  double z[2] = {0.0, 0.0};                                       // recurrence state of the two rows
  double scarry[2] = {0.0, 0.0};                                  // source function at the neighbouring level of the previous chunk
  double acc[2][8][2];
  double tacc[2][2];
  const int qbase = lane & ~3;

#pragma unroll 1
  for (int chunk = 0; chunk < n_chunk; ++chunk) {
    const int ci = up ? (n_chunk - 1 - chunk) : chunk;
    const int c0 = ci * SOS_CH;
    const int lb_al = max(c0 - 1, 0) & ~1;
#pragma unroll
    for (int mi = 0; mi < 2; ++mi)
#pragma unroll
      for (int ni = 0; ni < 8; ++ni) { acc[mi][ni][0] = 0.0; acc[mi][ni][1] = 0.0; }
    tacc[0][0] = tacc[0][1] = tacc[1][0] = tacc[1][1] = 0.0;
    const int nfr = (min(SOS_CH, L - c0) + 7) >> 3;

    for (int slab = 0; slab < n_slab; ++slab) {
      const int qi = chunk * n_slab + slab;
      const int stage = qi % STG;
      if (tid == 0) {
        if (qi + STG - 1 < n_iter) {
          if (qi >= 1) mbar_wait(empty + (qi - 1) % STG, ((qi - 1) / STG) & 1);
          issue(qi + STG - 1);
        }
        if (chunk >= 1 && slab == (n_slab >> 1)) {                 // tables of this chunk, once everybody left the previous one
          mbar_wait(tabfree, (chunk - 1) & 1);
          issue_tables(chunk);
        }
      }
      mbar_wait(full + stage, (qi / STG) & 1);
      if (own || LR) {
        const unsigned char *sp = smem_raw + stage * stage_bytes;
        const double *a = reinterpret_cast<const double *>(sp) + (wr * 16 + gq) * SOS_KB;
        const double *v = reinterpret_cast<const double *>(sp + STAGE_A_BYTES(rows_max)) + gq * SOS_KB;
        const double *b = reinterpret_cast<const double *>(sp + STAGE_A_BYTES(rows_max) + (LR ? STAGE_V_BYTES : 0)) + tq * SOS_SB + gq;
        const int kq = (slab * SOS_KB) % HB;
        const int ks_lim = (kq + SOS_KB <= 3 * N) ? 4 : max(0, (3 * N - kq + 3) >> 2);
        switch (nfr) {
          case 8: slab_mma<8, LR>(acc, tacc, a, v, b, own, wr, nw_launch, gq, tq, ks_lim); break;
          case 7: slab_mma<7, LR>(acc, tacc, a, v, b, own, wr, nw_launch, gq, tq, ks_lim); break;
          case 6: slab_mma<6, LR>(acc, tacc, a, v, b, own, wr, nw_launch, gq, tq, ks_lim); break;
          case 5: slab_mma<5, LR>(acc, tacc, a, v, b, own, wr, nw_launch, gq, tq, ks_lim); break;
          case 4: slab_mma<4, LR>(acc, tacc, a, v, b, own, wr, nw_launch, gq, tq, ks_lim); break;
          case 3: slab_mma<3, LR>(acc, tacc, a, v, b, own, wr, nw_launch, gq, tq, ks_lim); break;
          case 2: slab_mma<2, LR>(acc, tacc, a, v, b, own, wr, nw_launch, gq, tq, ks_lim); break;
          default: slab_mma<1, LR>(acc, tacc, a, v, b, own, wr, nw_launch, gq, tq, ks_lim); break;
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(empty + stage);
    }

    // ---------------- epilogue of this warp's 16 rows, in registers ----------------
    const double *sTc = sT + (chunk & 1) * 4 * SOS_CH;
    if (LR) {
      double *sTw = sT + (chunk & 1) * 4 * SOS_CH;
      if (gq < 4) {
#pragma unroll
        for (int u = 0; u < 2; ++u) {
          const int ni = wr + u * nw_launch;
          if (ni < 8) { sTw[gq * SOS_CH + ni * 8 + 2 * tq] = tacc[u][0]; sTw[gq * SOS_CH + ni * 8 + 2 * tq + 1] = tacc[u][1]; }
        }
      }
    }
    mbar_wait(tabfull, chunk & 1);
    const int off = c0 - lb_al;                                  // table row of column 0 (0 or 2)
    if (up && chunk == 0) {
      // The chunk that holds level NT (swept first): blank the table entries of the levels / layers that do not exist,
      // so that the element code below needs no level predicates: S = 0, a = 0, c = 0 there, and z stays 0 until
      // level NT, where c is replaced by the ground boundary value.
      const int cN = NT - c0;
      for (int c = cN + 1 + tid; c < SOS_CH; c += nthr) { sXd[c] = 0.0; sYd[c] = 0.0; }
      const int rs = NT - lb_al;
      for (int r = rs + tid; r < 72; r += nthr) { sDt[r] = 0.0; sInv[r] = 0.0; }
      if (ATTS) for (int idx = rs * N + tid; idx < 66 * N; idx += nthr) sAtt[idx] = 0.0;
    }
    if (LR || (up && chunk == 0)) __syncthreads();               // sT of this chunk / blanked tables visible (CTA-uniform)
    if (own) {
      const double *xdp = sXd + 2 * tq, *ydp = sYd + 2 * tq;
      // --- source function S = XDEL * (A_A X) + YDEL * (A_R X) in place (SOS_FSOURCE_ORDREIG) ---
#pragma unroll
      for (int ni = 0; ni < 8; ++ni)
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const double xd = xdp[ni * 8 + e];
#pragma unroll
          for (int mi = 0; mi < 2; ++mi) {
            double v = xd * acc[mi][ni][e];
            if (LR) {
              const int col = ni * 8 + 2 * tq + e;
              v = v + ydp[ni * 8 + e] * (u0[mi] * sTc[col] + us[mi] * sTc[(ty[mi] + 1) * SOS_CH + col]);
            }
            acc[mi][ni][e] = v;
          }
        }
      const int N8 = 8 * N;
      if (up) {
        // table row of column col is col + off; this thread's columns are ni*8 + 2*tq + e
        const double *dp = sDt + off + 2 * tq, *ip = sInv + off + 2 * tq;
        const double *abase = ATTS ? (sAtt + (off + 2 * tq) * N) : (tm.att + (size_t)(c0 + 2 * tq) * N);
        const double *ap0 = abase + (kidx[0] - 1), *ap1 = abase + (kidx[1] - 1);
        // --- layer constants c(i) = (1-a)(A mu + S(i)) - A a dtau, A = (S(i+1) - S(i)) / dtau  (SOS_OS.F:2279-2310) ---
        double edge[2];
#pragma unroll
        for (int mi = 0; mi < 2; ++mi) edge[mi] = __shfl_sync(0xffffffffu, acc[mi][0][0], qbase);   // S at column 0
#pragma unroll
        for (int ni = 0; ni < 8; ++ni) {
          const double dl0 = dp[ni * 8], dl1 = dp[ni * 8 + 1], iv0 = ip[ni * 8], iv1 = ip[ni * 8 + 1];
#pragma unroll
          for (int mi = 0; mi < 2; ++mi) {
            const double *ap = (mi ? ap1 : ap0) + ni * N8;
            const double nx_lane = __shfl_sync(0xffffffffu, acc[mi][ni][0], qbase | ((tq + 1) & 3));
            const double nx_blk = (ni < 7) ? __shfl_sync(0xffffffffu, acc[mi][ni < 7 ? ni + 1 : 7][0], qbase) : scarry[mi];
            const double s0 = acc[mi][ni][0], s1 = acc[mi][ni][1];
            const double s2 = (tq == 3) ? nx_blk : nx_lane;
            const double a0 = ap[0], a1 = ap[N];
            const double A0 = (s1 - s0) * iv0, A1 = (s2 - s1) * iv1;
            acc[mi][ni][0] = (1.0 - a0) * (A0 * mu_r[mi] + s0) - A0 * (a0 * dl0);
            acc[mi][ni][1] = (1.0 - a1) * (A1 * mu_r[mi] + s1) - A1 * (a1 * dl1);
          }
        }
        if (chunk == 0) {                                         // level NT: X = boundary value (z*0 + bc)
          const int cN = NT - c0;
#pragma unroll
          for (int ni = 0; ni < 8; ++ni)
#pragma unroll
            for (int e = 0; e < 2; ++e)
              if (ni * 8 + 2 * tq + e == cN) { acc[0][ni][e] = bc_r[0]; acc[1][ni][e] = bc_r[1]; }
        }
        // --- recurrence z <- z a + c from the ground upwards (descending level); a quad hands z from lane to lane ---
#pragma unroll
        for (int ni = 7; ni >= 0; --ni) {
          double aa[2][2], zin[2];
#pragma unroll
          for (int mi = 0; mi < 2; ++mi) {
            const double *ap = (mi ? ap1 : ap0) + ni * N8;
            aa[mi][0] = ap[0]; aa[mi][1] = ap[N];
            zin[mi] = z[mi];
          }
#pragma unroll
          for (int tt = 3; tt >= 0; --tt) {
#pragma unroll
            for (int mi = 0; mi < 2; ++mi) {
              if (tq == tt) zin[mi] = z[mi];
              double tv = z[mi] * aa[mi][1] + acc[mi][ni][1];
              tv = tv * aa[mi][0] + acc[mi][ni][0];
              z[mi] = __shfl_sync(0xffffffffu, tv, qbase | tt);
            }
          }
#pragma unroll
          for (int mi = 0; mi < 2; ++mi) {
            const double o1 = zin[mi] * aa[mi][1] + acc[mi][ni][1];
            acc[mi][ni][1] = o1;
            acc[mi][ni][0] = o1 * aa[mi][0] + acc[mi][ni][0];
          }
        }
#pragma unroll
        for (int mi = 0; mi < 2; ++mi) scarry[mi] = edge[mi];
      } else {
        // layer of level i is i - 1: table row of column col is col - 1 + off (column 0 of level 0 is clamped; its
        // constant is replaced by 0 below)
        const int rb = off + 2 * tq - 1;
        const double *dp = sDt + rb, *ip = sInv + rb;
        const double *abase = ATTS ? (sAtt + rb * N) : (tm.att + (ptrdiff_t)(c0 + 2 * tq - 1) * N);
        const double *ap0 = abase + (kidx[0] - 1), *ap1 = abase + (kidx[1] - 1);
        const int fix0 = (rb < 0) ? 1 : 0;                        // only (c0 = 0, tq = 0): first element uses row 0
        double edge[2];
#pragma unroll
        for (int mi = 0; mi < 2; ++mi) edge[mi] = __shfl_sync(0xffffffffu, acc[mi][7][1], qbase | 3);   // S at column 63
        const double rmuk0 = -mu_r[0], rmuk1 = -mu_r[1];
#pragma unroll
        for (int ni = 7; ni >= 0; --ni) {
          const int f = (ni == 0) ? fix0 : 0;
          const double dl0 = dp[ni * 8 + f], dl1 = dp[ni * 8 + 1], iv0 = ip[ni * 8 + f], iv1 = ip[ni * 8 + 1];
#pragma unroll
          for (int mi = 0; mi < 2; ++mi) {
            const double *ap = (mi ? ap1 : ap0) + ni * N8;
            const double pv_lane = __shfl_sync(0xffffffffu, acc[mi][ni][1], qbase | ((tq + 3) & 3));
            const double pv_blk = (ni > 0) ? __shfl_sync(0xffffffffu, acc[mi][ni > 0 ? ni - 1 : 0][1], qbase | 3) : scarry[mi];
            const double s1 = acc[mi][ni][0], s2 = acc[mi][ni][1];
            const double s0 = (tq == 0) ? pv_blk : pv_lane;
            const double a0 = ap[f * N], a1 = ap[N];
            const double rmuk = mi ? rmuk1 : rmuk0;
            const double A0 = (s1 - s0) * iv0, A1 = (s2 - s1) * iv1;
            acc[mi][ni][0] = (1.0 - a0) * (A0 * rmuk + s1) + A0 * (a0 * dl0);
            acc[mi][ni][1] = (1.0 - a1) * (A1 * rmuk + s2) + A1 * (a1 * dl1);
          }
        }
        if (c0 == 0 && tq == 0) { acc[0][0][0] = 0.0; acc[1][0][0] = 0.0; }   // level 0: X = 0 (z = 0 on entry)
        // --- recurrence from the top downwards (ascending level) ---
#pragma unroll
        for (int ni = 0; ni < 8; ++ni) {
          const int f = (ni == 0) ? fix0 : 0;
          double aa[2][2], zin[2];
#pragma unroll
          for (int mi = 0; mi < 2; ++mi) {
            const double *ap = (mi ? ap1 : ap0) + ni * N8;
            aa[mi][0] = ap[f * N]; aa[mi][1] = ap[N];
            zin[mi] = z[mi];
          }
#pragma unroll
          for (int tt = 0; tt < 4; ++tt) {
#pragma unroll
            for (int mi = 0; mi < 2; ++mi) {
              if (tq == tt) zin[mi] = z[mi];
              double tv = z[mi] * aa[mi][0] + acc[mi][ni][0];
              tv = tv * aa[mi][1] + acc[mi][ni][1];
              z[mi] = __shfl_sync(0xffffffffu, tv, qbase | tt);
            }
          }
#pragma unroll
          for (int mi = 0; mi < 2; ++mi) {
            const double o0 = zin[mi] * aa[mi][0] + acc[mi][ni][0];
            acc[mi][ni][0] = o0;
            acc[mi][ni][1] = o0 * aa[mi][1] + acc[mi][ni][1];
          }
        }
#pragma unroll
        for (int mi = 0; mi < 2; ++mi) scarry[mi] = edge[mi];
      }
      // --- new field, straight from the accumulator layout ---
      const int ncol = min(SOS_CH, L - c0);
#pragma unroll
      for (int mi = 0; mi < 2; ++mi) {
        if (!rvalid[mi]) continue;
        double *dst = xnext + SOS_XIDX(KP, r0 + wr * 16 + mi * 8 + gq, c0 + 2 * tq);
        if (ncol == SOS_CH) {
#pragma unroll
          for (int ni = 0; ni < 8; ++ni) *reinterpret_cast<double2 *>(dst + ni * 8) = make_double2(acc[mi][ni][0], acc[mi][ni][1]);
        } else {
#pragma unroll
          for (int ni = 0; ni < 8; ++ni) {
            const int col = ni * 8 + 2 * tq;
            if (col + 1 < ncol) *reinterpret_cast<double2 *>(dst + ni * 8) = make_double2(acc[mi][ni][0], acc[mi][ni][1]);
            else if (col < ncol) dst[ni * 8] = acc[mi][ni][0];
          }
        }
      }
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(tabfree);                         // this warp no longer reads the chunk's tables
  }
}

// =================================================================================================
// k_order1 -- first scattering order of every (term, Fourier order) item of a wave: the analytic source
// (SOS_FSOURCE_ORDRE1, SOS_OS.F:2431-2565, plus SOS_FSOURCE_DIFF_FRESNEL1 :3106-3295 for a flat sea; the layer
// integration is linear in the source, so both sources are integrated at once), the boundary values SOS_OS.F:970-992 and
// SOS_INTEGR_EPOPT (:2222-2357).  No contraction here: the kernel is bound by the 96*N*(NT+1)/2 bytes it writes per item.
// One thread per packed row sweeps all levels in the direction of propagation in blocks of 16 levels (two groups of 8:
// the layer constants of a group are independent, only the 8 recurrence FMAs are serial); a block goes through a small
// shared tile so that the CTA stores whole 128-byte runs per row.  Few registers, many CTAs per SM.
#define O1_PITCH 18
__global__ void __launch_bounds__(128)
k_order1(const ItemDev *__restrict__ items, const TermDev *__restrict__ terms, const OpticsDev *__restrict__ optics,
         const KsetDev *__restrict__ ksets)
{
  __shared__ __align__(16) double tile[128 * O1_PITCH];
  const ItemDev &it = items[blockIdx.x];
  const KsetDev &ks = ksets[it.kset];
  const TermDev &tm = terms[it.term];
  const OpticsDev &op = optics[tm.optics];
  const int N = op.nbmu, HB = op.HB, KP = op.KP, NT = tm.nt;
  const int tid = threadIdx.x;
  const int row = blockIdx.y * 128 + tid;
  const int dir = (row < KP) ? row / HB : 1, q = row - dir * HB;
  const bool valid = row < KP && q < 3 * N;                      // pad rows stay zero
  const bool up = (dir == 0);
  const int so = valid ? q / N : 0, kk = valid ? q % N + 1 : 1;
  const double mu = op.rmu[kk + N];
  const int rowc = valid ? row : 0;
  const double c1 = ks.c1[rowc], c2 = ks.c2[rowc], fz1 = ks.fz1[rowc], fz2 = ks.fz2[rowc];
  const bool fres = (op.ifresnel == 1);
  double *__restrict__ xo = it.x[1];
  const double *__restrict__ attp = tm.att + (kk - 1);
  const double *__restrict__ chp = tm.ch, *__restrict__ cfp = tm.cf, *__restrict__ xdp = tm.xdel, *__restrict__ ydp = tm.ydel;
  const double *__restrict__ dtp = tm.dt, *__restrict__ ivp = tm.inv;
  // source function of level lv (SOS_OS.F:2553-2560 and :3224-3292)
  auto source = [&](int lv) -> double {
    const double xdv = xdp[lv], ydv = ydp[lv];
    double v = __dmul_rn(chp[lv], __dadd_rn(__dmul_rn(c2, xdv), __dmul_rn(c1, ydv)));
    if (fres && (up ? (lv <= NT - 1) : (lv >= 1))) v = v + cfp[lv] * (fz2 * xdv + fz1 * ydv);
    return v;
  };
  double bc = 0.0;                                               // SOS_OS.F:970-992
  if (valid && up) {
    if (so == 0 && !(op.ro == 0.0 || it.is != 0)) bc = -op.ro * op.tab * tm.eground;
    if (op.imat_surf == 1) {
      const float *rs = op.surf + (size_t)it.is * 9 * N * N;
      const double rr = tm.eground / mu;
      double rv = (double)rs[(size_t)(so * 3) * N * N + (size_t)(kk - 1) * N + (op.n0 - 1)];
      if (op.ipolar == 0 && so != 0) rv = 0.0;
      bc = (so == 0) ? bc + rv * rr : rv * rr;
    }
  }
  const double rmuk = -mu;
  double z = 0.0, sedge = 0.0;
  double *trow = tile + tid * O1_PITCH;
  const int nblk = (NT + 16) >> 4;                               // blocks of 16 levels covering 0..NT
  for (int step = 0; step < nblk; ++step) {
    const int b16 = (up ? (nblk - 1 - step) : step) << 4;        // first level of this thread's block
    if (valid) {
#pragma unroll
      for (int g8 = 0; g8 < 2; ++g8) {
        const int b0 = b16 + (up ? 8 - 8 * g8 : 8 * g8);         // group of 8 levels, in the direction of propagation
        double *tg = trow + (b0 - b16);
        if (up) {
          if (b0 + 7 < NT) {                                     // whole group below the ground level
            double S[9], cst[8], aa[8], o[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) S[j] = source(b0 + j);
            S[8] = sedge;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const int l = b0 + j;
              const double a = attp[(size_t)l * N], dl = dtp[l], iv = ivp[l];
              const double A = (S[j + 1] - S[j]) * iv;
              cst[j] = (1.0 - a) * (A * mu + S[j]) - A * (a * dl);
              aa[j] = a;
            }
#pragma unroll
            for (int j = 7; j >= 0; --j) { z = z * aa[j] + cst[j]; o[j] = z; }
            sedge = S[0];
#pragma unroll
            for (int p = 0; p < 4; ++p) *reinterpret_cast<double2 *>(tg + 2 * p) = make_double2(o[2 * p], o[2 * p + 1]);
          } else {
            for (int lv = min(b0 + 7, NT); lv >= b0; --lv) {      // the group that holds level NT
              const double s = source(lv);
              if (lv == NT) z = bc;
              else {
                const double a = attp[(size_t)lv * N], dl = dtp[lv], iv = ivp[lv];
                const double A = (sedge - s) * iv;
                z = z * a + ((1.0 - a) * (A * mu + s) - A * (a * dl));
              }
              sedge = s;
              tg[lv - b0] = z;
            }
          }
        } else {
          if (b0 + 7 <= NT) {
            double S[9], cst[8], aa[8], o[8];
            S[0] = sedge;
#pragma unroll
            for (int j = 0; j < 8; ++j) S[j + 1] = source(b0 + j);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const int l = max(b0 + j - 1, 0);
              const double a = attp[(size_t)l * N], dl = dtp[l], iv = ivp[l];
              const double A = (S[j + 1] - S[j]) * iv;
              cst[j] = (1.0 - a) * (A * rmuk + S[j + 1]) + A * (a * dl);
              aa[j] = a;
            }
            if (b0 == 0) { cst[0] = 0.0; aa[0] = 0.0; }           // level 0: X = 0
#pragma unroll
            for (int j = 0; j < 8; ++j) { z = z * aa[j] + cst[j]; o[j] = z; }
            sedge = S[8];
#pragma unroll
            for (int p = 0; p < 4; ++p) *reinterpret_cast<double2 *>(tg + 2 * p) = make_double2(o[2 * p], o[2 * p + 1]);
          } else {
            for (int lv = b0; lv <= min(b0 + 7, NT); ++lv) {
              const double s = source(lv);
              if (lv == 0) z = 0.0;
              else {
                const double a = attp[(size_t)(lv - 1) * N], dl = dtp[lv - 1], iv = ivp[lv - 1];
                const double A = (s - sedge) * iv;
                z = z * a + ((1.0 - a) * (A * rmuk + s) + A * (a * dl));
              }
              sedge = s;
              tg[lv - b0] = z;
            }
          }
        }
      }
    }
    __syncthreads();
    // ---- copy-out: 8 lanes store one row's 16 levels (128 bytes), a warp 4 rows per instruction ----
    {
      const int part = tid & 7;
      for (int rl = tid >> 3; rl < 128; rl += 16) {
        const int r = blockIdx.y * 128 + rl;
        if (r >= KP) break;
        const int d = r / HB;
        if (r - d * HB >= 3 * N) continue;
        const int rb16 = ((d == 0) ? (nblk - 1 - step) : step) << 4;
        const int lv = rb16 + 2 * part;
        const double2 t = *reinterpret_cast<const double2 *>(tile + rl * O1_PITCH + 2 * part);
        double *dst = xo + SOS_XIDX(KP, r, lv);
        if (lv + 1 <= NT) *reinterpret_cast<double2 *>(dst) = t;
        else if (lv <= NT) *dst = t.x;
      }
    }
    __syncthreads();
  }
}

extern "C" int sos_launch_order1(const ItemDev *items, const TermDev *terms, const OpticsDev *optics, const KsetDev *ksets,
                                 int nitem, int maxKP, cudaStream_t st)
{
  if (nitem <= 0) return 0;
  const dim3 grid((unsigned)nitem, (unsigned)((maxKP + 127) / 128));
  k_order1<<<grid, 128, 0, st>>>(items, terms, optics, ksets);
  return 1;
}

static size_t step_smem_bytes(int nw, int lr, int order1, int att_cap)
{
  const size_t rows = (size_t)nw * 16;
  const size_t stage = STAGE_A_BYTES(rows) + (lr ? STAGE_V_BYTES : 0) + STAGE_B_BYTES;
  const size_t pipe = order1 ? 0 : SOS_STAGES * stage;
  const size_t sj = rows * SOS_SJ * 8;
  return (pipe > sj ? pipe : sj) + (4 * SOS_CH + 3 * 80) * 8 + (2 * SOS_STAGES + 2) * 8 + (6 * 72 + 128 + 4 * 128 + att_cap) * 8 + 128;
}

static size_t step2_smem_bytes(int nw, int lr, int att_cap, int stages)
{
  const size_t rows = (size_t)nw * 16;
  const size_t stage = STAGE_A_BYTES(rows) + (lr ? STAGE_V_BYTES : 0) + STAGE_B_BYTES;
  return stages * stage + (2 * 4 * SOS_CH + 3 * 80) * 8 + (2 * SOS_STAGES + 2) * 8 + (4 * 72 + 128 + 128 + att_cap) * 8 + 128;
}

extern "C" int sos_launch_step(const ItemDev *items, const TermDev *terms, const OpticsDev *optics,
                               const KsetDev *ksets, const int *list, int nitem, int order1, int mode,
                               int maxHB, double *jdump, cudaStream_t st)
{
  if (nitem <= 0) return 0;
  static bool attr_done = false;
  if (!attr_done) {
    const int big = 160 * 1024;
    cudaFuncSetAttribute(k_step<0, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, big);
    cudaFuncSetAttribute(k_step<1, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, big);
    cudaFuncSetAttribute(k_step<0, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, big);
    cudaFuncSetAttribute(k_step2<0, 3, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, big);
    cudaFuncSetAttribute(k_step2<1, 3, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, big);
    cudaFuncSetAttribute(k_step2<0, 4, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, big);
    cudaFuncSetAttribute(k_step2<1, 4, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, big);
    cudaFuncSetAttribute(k_step2<0, 7, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448);
    cudaFuncSetAttribute(k_step2<1, 7, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448);
    attr_done = true;
  }
  const int groups = maxHB / 16;
  const int tiles_per_dir = (groups + SOS_MAXW - 1) / SOS_MAXW;
  const int nw = std::max((groups + tiles_per_dir - 1) / tiles_per_dir, 4);   // >= 4 warps: T = V X needs 8 column blocks / 2
  const int nmax = maxHB / 3;                                    // 3N <= HB
  const int att_cap = (nmax <= 48) ? 66 * nmax : 0;            // stage the attenuation table of a chunk in smem when it fits
  static const int dbg = getenv("SOS_DBG") ? atoi(getenv("SOS_DBG")) : 0;   // 16: phase timing (SOS_PHASE_TIMING builds), 128: k_step2
  const dim3 grid((unsigned)nitem * 2 * tiles_per_dir);
  const dim3 block(nw * 32);
  int launches = 0;
  if (order1) {
    k_step<0, 1><<<grid, block, step_smem_bytes(nw, 0, 1, att_cap), st>>>(items, terms, optics, ksets, list, tiles_per_dir, 0, att_cap, jdump, dbg);
    launches = 1;
  } else if ((dbg & 128) && !jdump && att_cap > 0) {             // experimental: register-resident epilogue, see DESIGN.md 7
    const bool deep = (dbg & 256) != 0;                          // 4 stages, attenuation table read from global memory
    const bool solo = (dbg & 512) != 0;                          // one CTA per SM, 7 stages (untested on hardware: round 2)
    for (int lr = 0; lr < 2; ++lr) {
      if (!(mode & (1 << lr))) continue;
      if (solo) {
        // One CTA per SM with the whole shared memory as a deep pipeline (look-ahead 6 k-slabs) and up to 255
        // registers: tools/dmma_smem_bench shows that 8 warps saturate the DMMA pipe once operands are resident.
        if (lr) {
          const size_t sm = step2_smem_bytes(nw, 1, att_cap, 7);
          k_step2<1, 7, 1><<<grid, block, sm, st>>>(items, terms, optics, ksets, list, tiles_per_dir, 1, att_cap);
        } else {
          const size_t sm = step2_smem_bytes(nw, 0, att_cap, 7);
          k_step2<0, 7, 1><<<grid, block, sm, st>>>(items, terms, optics, ksets, list, tiles_per_dir, 0, att_cap);
        }
      } else if (deep) {
        const size_t sm = step2_smem_bytes(nw, lr, 0, 4);
        if (lr) k_step2<1, 4, 0><<<grid, block, sm, st>>>(items, terms, optics, ksets, list, tiles_per_dir, 1, 0);
        else k_step2<0, 4, 0><<<grid, block, sm, st>>>(items, terms, optics, ksets, list, tiles_per_dir, 0, 0);
      } else {
        const size_t sm = step2_smem_bytes(nw, lr, att_cap, 3);
        if (lr) k_step2<1, 3, 1><<<grid, block, sm, st>>>(items, terms, optics, ksets, list, tiles_per_dir, 1, att_cap);
        else k_step2<0, 3, 1><<<grid, block, sm, st>>>(items, terms, optics, ksets, list, tiles_per_dir, 0, att_cap);
      }
      ++launches;
    }
  } else {
    if (mode & 1) {                                              // aerosol-only Fourier orders (is > 2)
      k_step<0, 0><<<grid, block, step_smem_bytes(nw, 0, 0, att_cap), st>>>(items, terms, optics, ksets, list, tiles_per_dir, 0, att_cap, jdump, dbg);
      ++launches;
    }
    if (mode & 2) {                                              // is <= 2: molecular rank-4 part carried along
      k_step<1, 0><<<grid, block, step_smem_bytes(nw, 1, 0, att_cap), st>>>(items, terms, optics, ksets, list, tiles_per_dir, 1, att_cap, jdump, dbg);
      ++launches;
    }
  }
  return launches;
}
