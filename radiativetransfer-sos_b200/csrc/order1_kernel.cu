// order1_kernel.cu -- first scattering order of libsosgpu.so (sm_100a).  Orders n >= 2 are in sweep_kernel.cu.
#include "sosgpu_internal.h"
#include "sosgpu_async.cuh"
#include <math.h>

// k_order1_general -- first scattering order of every (term, Fourier order) item of a wave: the analytic source
// (SOS_FSOURCE_ORDRE1, SOS_OS.F:2431-2565, plus SOS_FSOURCE_DIFF_FRESNEL1 :3106-3295 for a flat sea; the layer
// integration is linear in the source, so both sources are integrated at once), the boundary values SOS_OS.F:970-992 and
// SOS_INTEGR_EPOPT (:2222-2357).  No contraction here: the kernel is bound by the 96*N*(NT+1)/2 bytes it writes per item.
// One thread per packed row sweeps all levels in the direction of propagation in blocks of 16 levels (two groups of 8:
// the layer constants of a group are independent, only the 8 recurrence FMAs are serial); a block goes through a small
// shared tile so that the CTA stores whole 128-byte runs per row.  Few registers, many CTAs per SM.
#define O1_PITCH 18
__global__ void __launch_bounds__(128)
k_order1_general(const ItemDev *__restrict__ items, const TermDev *__restrict__ terms, const OpticsDev *__restrict__ optics,
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
  const double *__restrict__ gp = tm.gco + (kk - 1), *__restrict__ bp = tm.bco + (kk - 1);
  const double *__restrict__ sxd = tm.sxd, *__restrict__ syd = tm.syd, *__restrict__ cfp = tm.cf;
  const double *__restrict__ xdp = tm.xdel, *__restrict__ ydp = tm.ydel;
  // Source function of level lv (SOS_OS.F:2553-2560 and :3224-3292): CH*(S2*XDEL + S1*YDEL) with CH*XDEL, CH*YDEL tabulated
  // per level (k_beam); vector FP64 is the scarce resource of this kernel (about 1/8 of the DMMA rate on B200), so the
  // source costs 2 instructions per level and the layer update 3 (tables of k_att, as in k_sweep).
  auto source = [&](int lv) -> double {
    double v = c2 * sxd[lv] + c1 * syd[lv];
    if (fres && (up ? (lv <= NT - 1) : (lv >= 1))) v = v + cfp[lv] * (fz2 * xdp[lv] + fz1 * ydp[lv]);
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
            double S[9], o[8], aa[8], bb[8], gg[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) {                         // all table loads of the group in flight before any use
              const size_t l = (size_t)(b0 + j) * N;
              aa[j] = attp[l]; bb[j] = bp[l]; gg[j] = gp[l];
            }
#pragma unroll
            for (int j = 0; j < 8; ++j) S[j] = source(b0 + j);
            S[8] = sedge;
#pragma unroll
            for (int j = 7; j >= 0; --j) {
              z = z * aa[j] + (bb[j] * S[j] + gg[j] * S[j + 1]);
              o[j] = z;
            }
            sedge = S[0];
#pragma unroll
            for (int p = 0; p < 4; ++p) *reinterpret_cast<double2 *>(tg + 2 * p) = make_double2(o[2 * p], o[2 * p + 1]);
          } else {
            for (int lv = min(b0 + 7, NT); lv >= b0; --lv) {      // the group that holds level NT
              const double s = source(lv);
              if (lv == NT) z = bc;
              else z = z * attp[(size_t)lv * N] + (bp[(size_t)lv * N] * s + gp[(size_t)lv * N] * sedge);
              sedge = s;
              tg[lv - b0] = z;
            }
          }
        } else {
          if (b0 + 7 <= NT) {
            double S[9], o[8], aa[8], bb[8], gg[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const size_t l = (size_t)max(b0 + j - 1, 0) * N;
              aa[j] = attp[l]; bb[j] = bp[l]; gg[j] = gp[l];
            }
            S[0] = sedge;
#pragma unroll
            for (int j = 0; j < 8; ++j) S[j + 1] = source(b0 + j);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              z = z * aa[j] + (bb[j] * S[j + 1] + gg[j] * S[j]);
              if (b0 + j == 0) z = 0.0;                           // level 0: X = 0
              o[j] = z;
            }
            sedge = S[8];
#pragma unroll
            for (int p = 0; p < 4; ++p) *reinterpret_cast<double2 *>(tg + 2 * p) = make_double2(o[2 * p], o[2 * p + 1]);
          } else {
            for (int lv = b0; lv <= min(b0 + 7, NT); ++lv) {
              const double s = source(lv);
              if (lv == 0) z = 0.0;
              else {
                const size_t l = (size_t)(lv - 1) * N;
                z = z * attp[l] + (bp[l] * s + gp[l] * sedge);
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

// k_order1 -- the same first order for waves without a flat-sea (Fresnel) source, which is every wave of the BASELINE
// configurations.  The source is c2(row)*sx(level) + c1(row)*sy(level), so the layer update folds into
//   X(i) = X(i -+ 1)*a + c2*T2 + c1*T1      (3 FP64 instructions per element; tables of k_att, see TermDev)
// with three [level][mu] tables per direction that do not depend on the Stokes component or the Fourier order.  A CTA is
// 128 rows of ONE direction; the 16-level x N block of each table is one contiguous run in HBM, fetched by one TMA bulk
// copy into a shared ring one block ahead of the recurrence, so no thread ever waits on a global load: vector
// FP64 is scarce on B200 (about 1/8 of the DMMA rate) and the only other cost is the field it writes, 96*N*(NT+1)/2 bytes
// per item, stored as whole 128-byte runs per row through a small shared tile.  Levels above NT read zero table rows.
#define O1_ST 2                                                   // table ring: 2 stages measured best (4 CTAs/SM at N=41; 3 stages: 3 CTAs, 11 % slower)
__global__ void __launch_bounds__(128)
k_order1(const ItemDev *__restrict__ items, const TermDev *__restrict__ terms, const OpticsDev *__restrict__ optics,
         const KsetDev *__restrict__ ksets, int nslice, int stage_doubles, int nstage)
{
  extern __shared__ __align__(128) unsigned char o1_smem[];
  double *tile = reinterpret_cast<double *>(o1_smem);                          // [128][O1_PITCH]
  double *stages = tile + 128 * O1_PITCH;                                      // [nstage][3][16*N]
  unsigned long long *full = reinterpret_cast<unsigned long long *>(stages + (size_t)nstage * stage_doubles);
  const ItemDev &it = items[blockIdx.x];
  const KsetDev &ks = ksets[it.kset];
  const TermDev &tm = terms[it.term];
  const OpticsDev &op = optics[tm.optics];
  const int N = op.nbmu, HB = op.HB, KP = op.KP, NT = tm.nt;
  const int dirn = blockIdx.y / nslice, q0 = (blockIdx.y - dirn * nslice) * 128;
  if (q0 >= 3 * N) return;                                       // slice of pad rows only (uniform for the CTA)
  const int tid = threadIdx.x;
  const bool up = (dirn == 0);
  const int q = q0 + tid;
  const bool valid = q < 3 * N;
  const int row = valid ? dirn * HB + q : 0;
  const int so = valid ? q / N : 0, k0 = valid ? q % N : 0;
  const double c1 = ks.c1[row], c2 = ks.c2[row];
  double bc = 0.0;                                               // SOS_OS.F:970-992
  if (valid && up) {
    if (so == 0 && !(op.ro == 0.0 || it.is != 0)) bc = -op.ro * op.tab * tm.eground;
    if (op.imat_surf == 1) {
      const float *rs = op.surf + (size_t)it.is * 9 * N * N;
      const double rr = tm.eground / op.rmu[k0 + 1 + N];
      double rv = (double)rs[(size_t)(so * 3) * N * N + (size_t)k0 * N + (op.n0 - 1)];
      if (op.ipolar == 0 && so != 0) rv = 0.0;
      bc = (so == 0) ? bc + rv * rr : rv * rr;
    }
  }
  const double *ta = up ? tm.att : tm.adn, *t1 = up ? tm.o1u1 : tm.o1d1, *t2 = up ? tm.o1u2 : tm.o1d2;
  const int nblk = (NT + 16) >> 4;                               // blocks of 16 levels covering 0..NT
  const int tab = 16 * N;                                        // doubles per table block
  auto fetch = [&](int step) {                                   // thread 0: the three table blocks of `step` into its stage
    const int s = step % nstage;
    const size_t off = (size_t)((up ? (nblk - 1 - step) : step) << 4) * N;
    double *dst = stages + (size_t)s * stage_doubles;
    mbar_expect_tx(full + s, 3u * (unsigned)tab * 8u);
    bulk_g2s(dst, ta + off, (unsigned)tab * 8u, full + s);
    bulk_g2s(dst + tab, t1 + off, (unsigned)tab * 8u, full + s);
    bulk_g2s(dst + 2 * tab, t2 + off, (unsigned)tab * 8u, full + s);
  };
  if (tid == 0) {
    for (int s = 0; s < nstage; ++s) mbar_init(full + s, 1);
    fence_proxy_async();
    for (int s = 0; s < nstage && s < nblk; ++s) fetch(s);
  }
  __syncthreads();
  double *__restrict__ xo = it.x[1];
  double z = 0.0;
  double *trow = tile + tid * O1_PITCH;
  for (int step = 0; step < nblk; ++step) {
    const int b16 = (up ? (nblk - 1 - step) : step) << 4;        // first level of the block
    const int s = step % nstage;
    mbar_wait(full + s, (unsigned)((step / nstage) & 1));
    if (valid) {
      const double *sa = stages + (size_t)s * stage_doubles + k0, *s1 = sa + tab, *s2 = s1 + tab;
#pragma unroll
      for (int g8 = 0; g8 < 2; ++g8) {                           // two groups of 8 levels, in the direction of propagation
        const int j0 = up ? 8 - 8 * g8 : 8 * g8;
        double o[8];
        if (up) {
#pragma unroll
          for (int j = 7; j >= 0; --j) {
            const int jj = (j0 + j) * N;
            z = z * sa[jj] + (c2 * s2[jj] + c1 * s1[jj]);
            if (b16 + j0 + j == NT) z = bc;                      // ground level (rows above it are zero in the tables)
            o[j] = z;
          }
        } else {
#pragma unroll
          for (int j = 0; j < 8; ++j) {                          // level 0: zero table row -> X = 0
            const int jj = (j0 + j) * N;
            z = z * sa[jj] + (c2 * s2[jj] + c1 * s1[jj]);
            o[j] = z;
          }
        }
#pragma unroll
        for (int p = 0; p < 4; ++p) *reinterpret_cast<double2 *>(trow + j0 + 2 * p) = make_double2(o[2 * p], o[2 * p + 1]);
      }
    }
    __syncthreads();                                             // tile complete; every thread is done with stage s
    if (tid == 0 && step + nstage < nblk) fetch(step + nstage);
    // ---- copy-out: 8 lanes store one row's 16 levels (128 bytes), a warp 4 rows per instruction ----
    {
      const int part = tid & 7;
      const int lv = b16 + 2 * part;
      for (int rl = tid >> 3; rl < 128; rl += 16) {
        if (q0 + rl >= 3 * N) break;
        const double2 t = *reinterpret_cast<const double2 *>(tile + rl * O1_PITCH + 2 * part);
        double *dst = xo + SOS_XIDX(KP, dirn * HB + q0 + rl, lv);
        if (lv + 1 <= NT) *reinterpret_cast<double2 *>(dst) = t;
        else if (lv <= NT) *dst = t.x;
      }
    }
    __syncthreads();
  }
}


extern "C" int sos_launch_order1(const ItemDev *items, const TermDev *terms, const OpticsDev *optics, const KsetDev *ksets,
                                 int nitem, int maxKP, int maxN, int any_fresnel, cudaStream_t st)
{
  if (nitem <= 0) return 0;
  if (any_fresnel) {
    const dim3 grid((unsigned)nitem, (unsigned)((maxKP + 127) / 128));
    k_order1_general<<<grid, 128, 0, st>>>(items, terms, optics, ksets);
    return 1;
  }
  const int nslice = (maxKP / 2 + 127) / 128, stage_doubles = 3 * 16 * maxN;
  const int nstage = O1_ST;
  const size_t smem = (size_t)(128 * O1_PITCH + nstage * stage_doubles) * sizeof(double) + nstage * sizeof(unsigned long long);
  static bool attr_set[64] = {};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return -1;
  if (dev >= 0 && dev < 64 && !attr_set[dev]) {
    if (cudaFuncSetAttribute(k_order1, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024) != cudaSuccess) return -1;
    attr_set[dev] = true;
  }
  if (smem > 200 * 1024) return -1;
  const dim3 grid((unsigned)nitem, (unsigned)(2 * nslice));
  k_order1<<<grid, 128, smem, st>>>(items, terms, optics, ksets, nslice, stage_doubles, nstage);
  return 1;
}
