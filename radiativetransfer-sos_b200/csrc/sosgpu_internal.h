// sosgpu_internal.h -- device data layout shared by the CUDA translation units of libsosgpu.so
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

// ------------------------------------------------------------------------------------------------
// Row / column numbering of the packed problem (DESIGN.md "Data layout in HBM")
//   N  = nbmu, W = 2N+1, HB = roundup(3N,16) rows per direction block, KP = 2*HB.
//   row r in [0,KP): d = r / HB (0: mu>0 "up", 1: mu<0 "down"), q = r % HB,
//                    valid iff q < 3N, stokes = q / N (0 I, 1 Q, 2 U), k = q % N + 1.
//   A field X(0:NT,-N:N) x {I,Q,U} of the reference is stored chunk-major as X[level/64][r][68] (SOS_XIDX);
//   pad rows / pad levels / pad columns stay zero.  LP = roundup(NT+1, 8) is only the pitch of test dumps.
// ------------------------------------------------------------------------------------------------
#define SOS_KB 16      // k-slab of the contraction (doubles)
#define SOS_CH 64      // level chunk (columns of the DMMA tile)
#define SOS_SB 68      // smem row stride of the B slab   (CH + 4)
// Field layout in HBM: chunk-major, padded rows -- X[(level>>6)][row][SOS_SB], element (row, level):
// the 16 k-rows x 64 levels a k-slab needs are one contiguous 16*SOS_SB*8-byte block = ONE TMA bulk copy that
// lands in shared memory with the conflict-free padded pitch.
#define SOS_XIDX(KP, row, level) ((((size_t)((level) >> 6)) * (size_t)(KP) + (size_t)(row)) * SOS_SB + ((level) & 63))
#define SOS_XSIZE(KP, L) ((size_t)(((L) + SOS_CH - 1) / SOS_CH) * (size_t)(KP) * SOS_SB)

struct OpticsDev {            // one per optics entry (device copy, arrays in a slab)
  int nbmu, W, HB, KP, os_nb, n0, imat_surf, ifresnel, ipolar, igmax;
  double tab;                 // TAB = -mu_s (SOS_OS.F:706-715)
  double ro, ron, ind_surf, zout;
  double beta2, gamma2, alpha2;      // Rayleigh coefficients (SOS_OS.F:678-699)
  double f11sun, f12sun;             // F11(0), F12(0) of SOS_MAT_FRESNEL_PLAN_REFL
  const double *rmu, *ga;            // [W] (rmu[N] = tab)
  const double *alpha, *beta, *gamma, *zeta;   // [os_nb+1]
  const double *f11, *f12, *f33;     // [N+1]
  const float *surf;                 // [nrec][9][N][N] or null
  int n_surf_rec;
};

struct TermDev {              // one per (wavelength, CKD term)
  int optics, group, nt, LP, iborm;
  int jout;                   // output level index J for zout != -1 (interp between J-1 and J), else -1
  double zz;                  // interpolation weight ZZ (SOS_OS.F:1520)
  double aik, eground;        // eground = exp(H(NT)/TAB)
  const double *h, *xdel, *ydel;      // [nt+1] truncation-adapted profile (SOS.F:523-543)
  const double *dt, *inv;             // [nt] layer optical thickness and reciprocal
  const double *ch;                   // [nt+1] exp(-H/(-TAB))/4  (SOS_OS.F:837-839)
  const double *cf;                   // [nt+1] flat-sea source attenuation (SOS_OS.F:3219,3278) or null
  const double *att;                  // [nt][N] a = exp(-dt/mu_k)
  const double *gco, *bco;            // [nt][N] layer-integration weights g = (1-a)*mu/dt - a and 1 - a - g (k_att)
  const double *sxd, *syd;            // [nt+1] CH*XDEL, CH*YDEL: the order-1 source is c2*sxd + c1*syd (k_beam)
  const double *pup, *qup, *pdn, *qdn;   // [nt+1][N] the same weights times XDEL of the two levels, per sweep direction (k_att)
  // First-order tables, indexed by the level they update (k_att): with sx = CH*XDEL, sy = CH*YDEL of the two levels of a layer,
  //   upward   X(i) = X(i+1)*att[i] + c2*o1u2[i] + c1*o1u1[i],   o1u2 = b*sx(i) + g*sx(i+1),  o1u1 likewise with sy
  //   downward X(i) = X(i-1)*adn[i] + c2*o1d2[i] + c1*o1d1[i],   adn[i] = att[i-1], o1d2 = b*sx(i) + g*sx(i-1)
  const double *o1u1, *o1u2, *adn, *o1d1, *o1d2;
  double *i4;                         // [6][2N] running Fourier sums I4,Q4,U4,I5,Q5,U5 (component order below)
};

// component order of the 6N "TOA/BOA" vectors (histories, sums): c = d*3N + stokes*N + (k-1)
struct ItemDev {              // one per (term, Fourier order) in the current wave
  int term, is, kset;         // kset: index of the (optics, is) kernel set of this wave
  int n;                      // current scattering order IG
  int active, reason;
  double *x[2];               // ping-pong fields [KP][LP]
  double *hist_a, *hist_d, *sum3;     // [6N]
  double *rii;                        // [3N] direct surface term at TOA (SOS_OS.F:1051-1084)
  double *sumout;                     // [2][6N] sums at the two output levels (zout != -1) or null
  double *riiout;                     // [2][3N]
};

struct KsetDev {              // kernel set of one (optics, Fourier order)
  int optics, is, dual;       // dual: Rayleigh part present (is <= 2)
  double beta0;
  double *basis;              // [3][(os_nb+2)][W]  PSL,RSL,TSL(-1:NB,-N:N)
  double *ker;                // [6][W*W] BP,GR,GT,ARR,ART,ATT   (j,k) at [(k+N)*W + (j+N)]
  double *xpl;                // [3][W] XPL,XRL,XTL
  double *apackA;             // [KP/16 slabs][KP rows][16] slab-major, k swizzled by 4*(row&3): 0.5*GA(j)*sign*element
  double *apackR;             // [KP][KP] row-major molecular part (dense; source of the rank-4 factors)
  double *vpack;              // [2 dirs][KP/16 slabs][8][16] swizzled: functionals V of the rank-4 molecular part
  double *urow;               // [KP] XPL/XRL/XTL(k) of each packed row (I/Q/U rows): J_R = T0 + urow*T_type
  double *c1, *c2;            // [KP] order-1 row coefficients (Rayleigh / aerosol)
  double *fz1, *fz2;          // [KP] flat-sea order-1 row coefficients
};

// launch helpers implemented in the .cu files
#ifdef __cplusplus
extern "C" {
#endif
void sos_launch_basis(const KsetDev *ksets, const OpticsDev *optics, int nkset, cudaStream_t st);
void sos_launch_kernels(const KsetDev *ksets, const OpticsDev *optics, int nkset, int maxW, cudaStream_t st);
void sos_launch_pack(const KsetDev *ksets, const OpticsDev *optics, int nkset, int maxKP, cudaStream_t st);   // 2 kernels
void sos_launch_att(const TermDev *terms, const OpticsDev *optics, int nterm, int max_elems, cudaStream_t st);
void sos_launch_init(ItemDev *items, const TermDev *terms, const OpticsDev *optics, int nitem, cudaStream_t st);
// list_cur == null: items 0..ncur-1; count_cur != null: number of valid list entries on the device (ncur = upper bound)
void sos_launch_test(ItemDev *items, const TermDev *terms, const OpticsDev *optics,
                     const int *list_cur, const int *count_cur, int ncur, int *list_next, int *count_next, cudaStream_t st);
void sos_launch_fourier(ItemDev *items, TermDev *terms, const OpticsDev *optics, int nterm,
                        const int *item_of, int s0, int s1, int rec_stride_dev, int wdev,
                        double *rec, int *n_fourier, int *n_scatter, int *stop_reason,
                        double *emoins, double *eplus, int *done, cudaStream_t st);
void sos_launch_aggregate(const TermDev *terms, const int *group_start, const int *group_terms, int ngroup,
                          const double *rec, const int *n_fourier, int rec_stride_dev, int wdev,
                          double *grec, int *gnrec, cudaStream_t st);
// first scattering order of every item of a wave (analytic source, boundary values, layer integration); any_fresnel selects
// the general kernel that also integrates the flat-sea source.  Returns the launches made, -1 on a CUDA error.
int  sos_launch_order1(const ItemDev *items, const TermDev *terms, const OpticsDev *optics, const KsetDev *ksets,
                       int nitem, int maxKP, int maxN, int any_fresnel, cudaStream_t st);
// one scattering order n >= 2 (sweep_kernel.cu): persistent warp-specialised kernel; the item count is read from
// count_ptr on the device when non-null (nitem is then an upper bound)
int  sos_launch_sweep(const ItemDev *items, const TermDev *terms, const OpticsDev *optics, const KsetDev *ksets,
                      const int *list, const int *count_ptr, int nitem, int maxHB, unsigned *work_counter, int num_sms,
                      double *jdump, cudaStream_t st);
#ifdef __cplusplus
}
#endif
