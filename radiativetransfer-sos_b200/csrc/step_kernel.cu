// step_kernel.cu -- the fused hot kernel of libsosgpu.so (sm_100a).
//
// One launch advances every active (term, Fourier order) item by one scattering order:
//   J(level, mu) = XDEL(level) * (A_A X_{n-1})(level) + YDEL(level) * (A_R X_{n-1})(level)
//        = SOS_FSOURCE_ORDREIG (SOS_OS.F:2663-3017) as a dense FP64 contraction on DMMA tiles
//          (mma.sync.m8n8k4.f64), operands staged through shared memory by a 3-stage cp.async pipeline;
//   X_n = SOS_INTEGR_EPOPT(J, boundary values)  (SOS_OS.F:2222-2357), run on the tile while it is still
//          on chip: the source function never goes to HBM;
//   boundary values at the ground for mu>0 (Lambert / BRDF-BPDF quadrature / flat Fresnel,
//          SOS_OS.F:1166-1239) are formed in the prologue from the previous order's downward field.
// ORDER1 variant: the source is the analytic first-order source (SOS_FSOURCE_ORDRE1, SOS_OS.F:2431-2565,
//          plus SOS_FSOURCE_DIFF_FRESNEL1 for a flat sea) and the boundary values are SOS_OS.F:970-992.
//
// CTA tile: up to 128 rows (8 warps x 16 rows) of one direction block x 64 levels; the level chunks are
// swept in the direction of propagation so the recurrence state stays in registers.
#include "sosgpu_internal.h"
#include <math.h>

__device__ __forceinline__ void cp_async16(void *smem_dst, const void *gsrc, bool pred)
{
  const unsigned s = (unsigned)__cvta_generic_to_shared(smem_dst);
  const int sz = pred ? 16 : 0;
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(s), "l"(gsrc), "r"(sz));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int NKEEP>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(NKEEP)); }

__device__ __forceinline__ void dmma8x8x4(double &c0, double &c1, double a, double b)
{
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(c0), "+d"(c1)
               : "d"(a), "d"(b));
}

struct StepSmem {
  double *sA, *sA2, *sB, *sJ, *sG;
};

template <int DUAL, int ORDER1>
__global__ void __launch_bounds__(256, 1)
k_step(const ItemDev *__restrict__ items, const TermDev *__restrict__ terms, const OpticsDev *__restrict__ optics,
       const KsetDev *__restrict__ ksets, const int *__restrict__ list, int tiles_per_dir, int want_dual,
       double *__restrict__ jdump)
{
  extern __shared__ __align__(16) double smem[];
  const int nw_launch = blockDim.x >> 5;
  const int rows_max = nw_launch * 16;
  // carve shared memory
  double *sA = smem;
  double *sA2 = sA + (ORDER1 ? 0 : SOS_STAGES * rows_max * SOS_SA);
  double *sB = sA2 + ((DUAL && !ORDER1) ? SOS_STAGES * rows_max * SOS_SA : 0);
  double *sJ = sB + (ORDER1 ? 0 : SOS_STAGES * SOS_KB * SOS_SB);
  double *sG = sJ + rows_max * SOS_SJ;

  const int per_item = 2 * tiles_per_dir;
  const int ii = blockIdx.x / per_item, t = blockIdx.x % per_item;
  const int item = list ? list[ii] : ii;
  const ItemDev it = items[item];
  const KsetDev ks = ksets[it.kset];
  if (!ORDER1 && (ks.dual != want_dual)) return;               // handled by the other instantiation
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
  const bool up = (dir == 0);

  const double *__restrict__ xprev = it.x[it.n & 1];
  double *__restrict__ xnext = ORDER1 ? it.x[1] : it.x[(it.n + 1) & 1];

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
      for (int c = tid; c < 3 * N; c += blockDim.x) sG[c] = xprev[(size_t)(HB + c) * LP + NT];
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
          const double gi = sG[kk - 1], gq = sG[N + kk - 1], gu = sG[2 * N + kk - 1];
          if (so == 0) bc = bc + op.f11[kk] * gi + op.f12[kk] * gq;
          else if (so == 1) bc = bc + op.f12[kk] * gi + op.f11[kk] * gq;
          else bc = bc + op.f33[kk] * gu;
        }
      }
    }
  }

  // ---------------- level chunks, swept in the direction of propagation ----------------
  const int n_chunk = (L + SOS_CH - 1) / SOS_CH;
  const int n_slab = KP / SOS_KB;
  const int T = ORDER1 ? n_chunk : n_chunk * n_slab;
  const double *__restrict__ Ag = ks.apackA + (size_t)r0 * KP;
  const double *__restrict__ Ag2 = DUAL ? ks.apackR + (size_t)r0 * KP : nullptr;

  double acc[2][8][2];
  double acc2[DUAL ? 2 : 1][DUAL ? 8 : 1][2];
  double z = 0.0, sprev = 0.0;                                    // recurrence state of this thread's row

  auto issue = [&](int tt) {
    const int stage = tt % SOS_STAGES;
    const int chunk = tt / n_slab, slab = tt % n_slab;
    const int ci = up ? (n_chunk - 1 - chunk) : chunk;
    const int c0 = ci * SOS_CH, k0 = slab * SOS_KB;
    double *a = sA + stage * rows_max * SOS_SA;
    for (int idx = tid; idx < R * (SOS_KB / 2); idx += blockDim.x) {
      const int row = idx >> 3, c = idx & 7;
      cp_async16(a + row * SOS_SA + c * 2, Ag + (size_t)row * KP + k0 + c * 2, true);
    }
    if (DUAL) {
      double *a2 = sA2 + stage * rows_max * SOS_SA;
      for (int idx = tid; idx < R * (SOS_KB / 2); idx += blockDim.x) {
        const int row = idx >> 3, c = idx & 7;
        cp_async16(a2 + row * SOS_SA + c * 2, Ag2 + (size_t)row * KP + k0 + c * 2, true);
      }
    }
    double *b = sB + stage * SOS_KB * SOS_SB;
    for (int idx = tid; idx < SOS_KB * (SOS_CH / 2); idx += blockDim.x) {
      const int krow = idx >> 5, c = idx & 31;
      const int col = c0 + 2 * c;
      const bool ok = col < LP;
      cp_async16(b + krow * SOS_SB + 2 * c, ok ? (xprev + (size_t)(k0 + krow) * LP + col) : xprev, ok);
    }
  };

  if (!ORDER1) {
    for (int tt = 0; tt < SOS_STAGES - 1; ++tt) {
      if (tt < T) issue(tt);
      cp_async_commit();
    }
  }

  for (int tt = 0; tt < T; ++tt) {
    const int chunk = ORDER1 ? tt : tt / n_slab;
    const int slab = ORDER1 ? 0 : tt % n_slab;
    const int ci = up ? (n_chunk - 1 - chunk) : chunk;
    const int c0 = ci * SOS_CH;
    bool chunk_done = true;

    if (!ORDER1) {
      cp_async_wait<SOS_STAGES - 2>();
      __syncthreads();
      if (tt + SOS_STAGES - 1 < T) issue(tt + SOS_STAGES - 1);
      cp_async_commit();

      if (slab == 0) {
#pragma unroll
        for (int mi = 0; mi < 2; ++mi)
#pragma unroll
          for (int ni = 0; ni < 8; ++ni) {
            acc[mi][ni][0] = 0.0; acc[mi][ni][1] = 0.0;
            if (DUAL) { acc2[mi][ni][0] = 0.0; acc2[mi][ni][1] = 0.0; }
          }
      }
      if (wr < ng) {
        const int stage = tt % SOS_STAGES;
        const double *a = sA + stage * rows_max * SOS_SA + (wr * 16 + (lane >> 2)) * SOS_SA + (lane & 3);
        const double *a2 = sA2 + stage * rows_max * SOS_SA + (wr * 16 + (lane >> 2)) * SOS_SA + (lane & 3);
        const double *b = sB + stage * SOS_KB * SOS_SB + (lane & 3) * SOS_SB + (lane >> 2);
#pragma unroll
        for (int ks4 = 0; ks4 < SOS_KB / 4; ++ks4) {
          const double a0 = a[ks4 * 4], a1 = a[8 * SOS_SA + ks4 * 4];
          double ar0 = 0.0, ar1 = 0.0;
          if (DUAL) { ar0 = a2[ks4 * 4]; ar1 = a2[8 * SOS_SA + ks4 * 4]; }
#pragma unroll
          for (int ni = 0; ni < 8; ++ni) {
            const double bv = b[ks4 * 4 * SOS_SB + ni * 8];
            dmma8x8x4(acc[0][ni][0], acc[0][ni][1], a0, bv);
            dmma8x8x4(acc[1][ni][0], acc[1][ni][1], a1, bv);
            if (DUAL) {
              dmma8x8x4(acc2[0][ni][0], acc2[0][ni][1], ar0, bv);
              dmma8x8x4(acc2[1][ni][0], acc2[1][ni][1], ar1, bv);
            }
          }
        }
      }
      chunk_done = (slab == n_slab - 1);
    }
    if (!chunk_done) continue;

    // ---------------- tile epilogue: source function -> staging tile ----------------
    if (ORDER1) {
      __syncthreads();                                           // previous chunk's write-out has finished
      for (int idx = tid; idx < R * SOS_CH; idx += blockDim.x) {
        const int rowl = idx >> 6, col = idx & 63;
        const int level = c0 + col, row = r0 + rowl;
        double v = 0.0;
        if (level < L) {
          // SOS_FSOURCE_ORDRE1 (SOS_OS.F:2553-2560): ATTDIR*(S2*PCAER + S1*PCRAY)
          v = __dmul_rn(tm.ch[level], __dadd_rn(__dmul_rn(ks.c2[row], tm.xdel[level]), __dmul_rn(ks.c1[row], tm.ydel[level])));
          if (op.ifresnel == 1) {                                // SOS_FSOURCE_DIFF_FRESNEL1 (SOS_OS.F:3224-3292)
            const bool on = up ? (level <= NT - 1) : (level >= 1);
            if (on) v = v + tm.cf[level] * (ks.fz2[row] * tm.xdel[level] + ks.fz1[row] * tm.ydel[level]);
          }
        }
        sJ[rowl * SOS_SJ + col] = v;
      }
    } else {
      if (wr < ng) {
#pragma unroll
        for (int ni = 0; ni < 8; ++ni) {
#pragma unroll
          for (int e = 0; e < 2; ++e) {
            const int col = ni * 8 + 2 * (lane & 3) + e;
            const int level = c0 + col;
            double xd = 0.0, yd = 0.0;
            if (level < L) { xd = __ldg(tm.xdel + level); if (DUAL) yd = __ldg(tm.ydel + level); }
#pragma unroll
            for (int mi = 0; mi < 2; ++mi) {
              double v = xd * acc[mi][ni][e];
              if (DUAL) v = v + yd * acc2[mi][ni][e];
              sJ[(wr * 16 + mi * 8 + (lane >> 2)) * SOS_SJ + col] = v;
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

    // ---------------- SOS_INTEGR_EPOPT on the tile: one thread per row, sequential in level ----------------
    if (rowvalid) {
      double *js = sJ + tid * SOS_SJ;
      const double *att = tm.att + (kk - 1);
      if (up) {                                                  // SOS_OS.F:2279-2310
        const int hi = min(c0 + SOS_CH - 1, NT);
        for (int level = hi; level >= c0; --level) {
          const int col = level - c0;
          const double s = js[col];
          if (level == NT) {
            z = bc;
          } else {
            const double a = __ldg(att + (size_t)level * N);
            const double dl = __ldg(tm.dt + level), iv = __ldg(tm.inv + level);
            const double A = (sprev - s) * iv;
            z = z * a + (1.0 - a) * (A * mu + s) - A * (a * dl);
          }
          js[col] = z;
          sprev = s;
        }
      } else {                                                   // SOS_OS.F:2320-2354
        const int hi = min(c0 + SOS_CH - 1, NT);
        const double rmuk = -mu;
        for (int level = c0; level <= hi; ++level) {
          const int col = level - c0;
          const double s = js[col];
          if (level == 0) {
            z = 0.0;
          } else {
            const double a = __ldg(att + (size_t)(level - 1) * N);
            const double dl = __ldg(tm.dt + level - 1), iv = __ldg(tm.inv + level - 1);
            const double A = (s - sprev) * iv;
            z = z * a + (1.0 - a) * (A * rmuk + s) + A * (a * dl);
          }
          js[col] = z;
          sprev = s;
        }
      }
    }
    __syncthreads();

    // ---------------- write the new field tile (coalesced along levels) ----------------
    for (int idx = tid; idx < R * SOS_CH; idx += blockDim.x) {
      const int rowl = idx >> 6, col = idx & 63;
      const int level = c0 + col, row = r0 + rowl;
      if (level < L && (row - dir * HB) < 3 * N) xnext[(size_t)row * LP + level] = sJ[rowl * SOS_SJ + col];
    }
  }
}

static size_t step_smem_bytes(int nw, int dual, int order1)
{
  const size_t rows = (size_t)nw * 16;
  size_t d = 0;
  if (!order1) {
    d += (size_t)SOS_STAGES * rows * SOS_SA * (dual ? 2 : 1);
    d += (size_t)SOS_STAGES * SOS_KB * SOS_SB;
  }
  d += rows * SOS_SJ;
  d += 3 * 80;
  return d * sizeof(double);
}

extern "C" int sos_launch_step(const ItemDev *items, const TermDev *terms, const OpticsDev *optics,
                               const KsetDev *ksets, const int *list, int nitem, int order1, int mode,
                               int maxHB, double *jdump, cudaStream_t st)
{
  if (nitem <= 0) return 0;
  static bool attr_done = false;
  if (!attr_done) {
    const int big = 227 * 1024;
    cudaFuncSetAttribute(k_step<0, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, big);
    cudaFuncSetAttribute(k_step<1, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, big);
    cudaFuncSetAttribute(k_step<0, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, big);
    attr_done = true;
  }
  const int groups = maxHB / 16;
  const int tiles_per_dir = (groups + SOS_MAXW - 1) / SOS_MAXW;
  const int nw = (groups + tiles_per_dir - 1) / tiles_per_dir;
  const dim3 grid((unsigned)nitem * 2 * tiles_per_dir);
  const dim3 block(nw * 32);
  int launches = 0;
  if (order1) {
    k_step<0, 1><<<grid, block, step_smem_bytes(nw, 0, 1), st>>>(items, terms, optics, ksets, list, tiles_per_dir, 0, jdump);
    launches = 1;
  } else {
    // items whose Fourier order carries a Rayleigh part (is <= 2) run the dual-accumulator instantiation
    if (mode & 1) {
      k_step<0, 0><<<grid, block, step_smem_bytes(nw, 0, 0), st>>>(items, terms, optics, ksets, list, tiles_per_dir, 0, jdump);
      ++launches;
    }
    if (mode & 2) {
      k_step<1, 0><<<grid, block, step_smem_bytes(nw, 1, 0), st>>>(items, terms, optics, ksets, list, tiles_per_dir, 1, jdump);
      ++launches;
    }
  }
  return launches;
}
