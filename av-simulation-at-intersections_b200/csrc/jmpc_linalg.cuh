// jmpc_linalg.cuh -- warp-cooperative dense linear algebra on 4x4 tiles in shared memory (fp64).
//
// A symmetric n x n matrix (n = 2T <= 62) is stored as its lower block triangle of 4x4 tiles: tile (I, J),
// I >= J, sits at slot I(I+1)/2 + J, each slot kTS = 18 doubles (16 used, row major).  The 144-byte slot stride
// keeps every tile 16-byte aligned for 128-bit shared loads and spreads tiles that different lanes touch at the
// same in-tile offset over different banks (36 words apart; 128-bit accesses are served per quarter warp).
// Diagonal tiles of a *matrix* (not of a factor) are stored full (both triangles).  n is padded to a multiple of
// 4 with identity rows.
//
// None of the shared-memory pointers below is __restrict__: lanes exchange data through them around
// __syncwarp(), and a noalias parameter lets the compiler move its loads/stores across that call (it did, as soon
// as the block count became a compile-time constant and the loops were unrolled).
//
// Compared with a column-at-a-time factorisation on a packed triangle this does 64 FMAs per 32 shared-memory
// instructions in the trailing update instead of 1 per 2, and synchronises the warp 3 times per 4 columns.
//
// Lane groups: every routine is a template on G, the number of lanes that cooperate on one matrix (32 = the whole
// warp, 16 = two independent matrices per warp, one per half).  `gl` is the lane's index inside its group, `gm` the
// group's lane mask; synchronisation and shuffles are group-scoped, so the halves of a warp never depend on each
// other's control flow (they run in lock step whenever their control flow agrees, which is the common case).
#pragma once
#include <cuda_runtime.h>
#include <math.h>

namespace jmpc {

// The block-column loops of the factorisation, the triangular sweeps and the matvec stay rolled: unrolled (tile offsets as
// immediates) the interior-point loop is 37-55 KB of SASS, rolled 27-30 KB, and the instruction cache behind the
// schedulers holds 32 KB -- measured on the B200: T = 8 15.8 -> 21.7 M solves/s, T = 13 / 20 / 25 +2-4 % on top of
// the shorter row work.  -DJMPC_UNROLL_BLOCKS brings the unrolled loops back for comparison.
#ifdef JMPC_UNROLL_BLOCKS
#define JMPC_PRAGMA_CHOL _Pragma("unroll")
#define JMPC_PRAGMA_SWEEP _Pragma("unroll")
#define JMPC_PRAGMA_SYMV _Pragma("unroll")
#else
#define JMPC_PRAGMA_CHOL _Pragma("unroll 1")
#define JMPC_PRAGMA_SWEEP _Pragma("unroll 1")
#define JMPC_PRAGMA_SYMV _Pragma("unroll 1")
#endif

// Cycle accounting of a debug build (-DJMPC_CYCLES): lane 0 accumulates clock64() deltas per code region into
// g_cycles[]; read back with jmpc_debug_cycles.  Meant for single-instance runs (one warp alone on the GPU), where it
// gives the latency breakdown of the critical path.
#ifdef JMPC_CYCLES
__device__ unsigned long long g_cycles[32];
#define JMPC_TICK(var) long long var = clock64()
#define JMPC_TOCK(var, slot) do { long long now_ = clock64(); if ((threadIdx.x & 31) == 0) atomicAdd(&g_cycles[slot], (unsigned long long)(now_ - var)); var = now_; } while (0)
#else
#define JMPC_TICK(var)
#define JMPC_TOCK(var, slot)
#endif

constexpr int kTS = 18;                               // doubles per tile slot
constexpr unsigned kFullMask = 0xffffffffu;

__host__ __device__ inline int nblk(int n) { return (n + 3) >> 2; }
__host__ __device__ inline int tiles_doubles(int n) { const int nb = nblk(n); return (nb * (nb + 1) / 2) * kTS; }
__host__ __device__ __forceinline__ int tile_off(int I, int J) { return (((I * (I + 1)) >> 1) + J) * kTS; }
// element (i, j) with j <= i, or any (i, j) inside a diagonal tile
__host__ __device__ __forceinline__ int elem_off(int i, int j) { return tile_off(i >> 2, j >> 2) + ((i & 3) << 2) + (j & 3); }

// 1 / sqrt(a) for a > 0 in the normal range (pivots here are 1e-2 .. 1e20): the hardware seed (MUFU.RSQ64H, 2^-23) and
// one cubic correction, as the library's fast path, without its exponent-range test and call.  The library version
// put two branches and a reconvergence point into each of the four pivots of a diagonal block.
__device__ __forceinline__ double rsqrt_pos(double a) {
  double y;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(a));
  const double e = fma(-a, y * y, 1.0);
  return fma(fma(e, 0.375, 0.5), y * e, y);
}
// 1 / a for a > 0 in the normal range (slacks: 1e-30 .. 1e6): hardware seed (MUFU.RCP64H) and two Newton steps, without the
// library's exponent-range test and slow-path call.  Not correctly rounded (1 ulp); the interior-point iteration only
// needs s * (1/s) = 1 to working precision.
__device__ __forceinline__ double rcp_pos(double a) {
  double y;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(a));
  double e = fma(-a, y, 1.0);
  y = fma(y, e, y);
  e = fma(-a, y, 1.0);
  return fma(y, e, y);
}
// reciprocal root of a pivot; a non-positive pivot is treated as infinite (see chol_tiles)
__device__ __forceinline__ double pivot_rsqrt(double a, bool& clean) {
  const double y = rsqrt_pos(a);              // NaN for a <= 0, discarded by the select
  const bool pos = a > 0.0;
  clean = clean && pos;
  return pos ? y : 0.0;
}

__device__ __forceinline__ void ld4(const double* p, double& a, double& b, double& c, double& d) {
  const double2 x = *reinterpret_cast<const double2*>(p), y = *reinterpret_cast<const double2*>(p + 2);
  a = x.x; b = x.y; c = y.x; d = y.y;
}
__device__ __forceinline__ void st4(double* p, double a, double b, double c, double d) {
  *reinterpret_cast<double2*>(p) = make_double2(a, b);
  *reinterpret_cast<double2*>(p + 2) = make_double2(c, d);
}

// Task table of the trailing update: for block column J, task q (half a tile) -> tile row I = J + 1 + a, tile
// column Kc = J + 1 + b (b <= a), half h.  (a, b, h) only depend on m = nb - J - 1; the entry carries the slot
// numbers of the two factor tiles the task reads for the launch's block count nb (J = nb - 1 - m), so the kernel
// forms its three addresses with one multiply-add each instead of evaluating tile_off() three times (the integer
// work was a quarter of the update's instructions).  One table of sum_{m=1}^{nb-1} m (m + 1) 32-bit entries per
// block, built once per launch in shared memory.  (Decoding q arithmetically -- float sqrt plus fix-up loops -- was
// 4 % of the kernel's instructions.)
//   bits 0-7   slot of tile (I, J)            bits 8-15  slot of tile (Kc, J)
//   bits 16-19 b + 1 (slot of (I, Kc) = slot of (I, J) + b + 1)       bit 22  h (so (e >> 16) & 0x40 = byte offset of the half)
typedef unsigned int chol_task;
__host__ __device__ inline int chol_lut_entries(int nb) { return ((nb - 1) * nb * (nb + 1)) / 3; }
__host__ __device__ __forceinline__ int chol_lut_offset(int m) { return ((m - 1) * m * (m + 1)) / 3; }   // tasks of all m' < m
__host__ __device__ inline size_t chol_lut_bytes(int nb) { return ((size_t)chol_lut_entries(nb) * sizeof(chol_task) + 15) & ~(size_t)15; }
// table for a factorisation with nb block rows (rows m = 1 .. nb - 1)
__device__ inline void chol_lut_build(chol_task* lut, int nb, int tid, int nthreads) {
  for (int m = 1; m < nb; ++m) {
    const int ntasks = m * (m + 1), off = chol_lut_offset(m);
    const int J = nb - 1 - m;
    for (int q = tid; q < ntasks; q += nthreads) {
      const int t = q >> 1;
      int a = (int)((sqrtf(8.0f * (float)t + 1.0f) - 1.0f) * 0.5f);
      while (((a + 1) * (a + 2) >> 1) <= t) ++a;
      while (((a * (a + 1)) >> 1) > t) --a;
      const int b = t - ((a * (a + 1)) >> 1);
      const int I = J + 1 + a, Kc = J + 1 + b;
      lut[off + q] = ((unsigned)(b + 1) << 16) | ((unsigned)(q & 1) << 22) | (unsigned)(((I * (I + 1)) >> 1) + J) |
                     ((unsigned)(((Kc * (Kc + 1)) >> 1) + J) << 8);
    }
  }
}

// In-place Cholesky K = L L' on tiles.  The off-diagonal tiles receive L; each diagonal tile receives the INVERSE of
// L's diagonal block (lower triangular), which turns the panel and the triangular solves into multiplications.
// Returns false when some pivot was not positive and had to be replaced (all lanes agree); the factor is usable
// either way.
// When `rhs` is given (4 * nb doubles in shared memory) the forward substitution L y = rhs rides along: block J of y
// is the inverted diagonal block times block J of rhs (every lane, registers), and the lane that has just computed a
// row of the panel subtracts that row times y from its rhs entry -- no extra loads of L and no extra
// synchronisation; rhs holds y afterwards (solve_backward_tiles completes the solve).
template <int G>
__device__ inline bool chol_tiles(double* K, int nb, int gl, unsigned gm, const chol_task* lut,
                                  double* rhs = nullptr) {
  bool all_clean = true;
  JMPC_PRAGMA_CHOL
  for (int J = 0; J < nb; ++J) {
    JMPC_TICK(tc_);
    // ---- diagonal block: every lane factors it redundantly in registers (no broadcast needed)
    const double* D = K + tile_off(J, J);
    double a00, a10, a11, a20, a21, a22, a30, a31, a32, a33;
    {
      double u01, u02, u03, u12, u13, u23;             // upper triangle of the block: loaded, not used
      ld4(D, a00, u01, u02, u03); ld4(D + 4, a10, a11, u12, u13);
      ld4(D + 8, a20, a21, a22, u23); ld4(D + 12, a30, a31, a32, a33);
      (void)u01; (void)u02; (void)u03; (void)u12; (void)u13; (void)u23;
    }
    // A pivot that is not positive (roundoff once the barrier weights reach ~1e13) is treated as infinite, the
    // usual interior-point remedy: its reciprocal root is set to 0, which zeroes the column of L and that
    // component of every solve; the outer iteration corrects the step.  `clean` reports whether it happened.
    bool clean = true;
    const double r0 = pivot_rsqrt(a00, clean);
    const double l10 = a10 * r0, l20 = a20 * r0, l30 = a30 * r0;
    a11 = fma(-l10, l10, a11);
    const double r1 = pivot_rsqrt(a11, clean);
    const double l21 = fma(-l20, l10, a21) * r1, l31 = fma(-l30, l10, a31) * r1;
    a22 = fma(-l21, l21, fma(-l20, l20, a22));
    const double r2 = pivot_rsqrt(a22, clean);
    const double l32 = fma(-l31, l21, fma(-l30, l20, a32)) * r2;
    a33 = fma(-l32, l32, fma(-l31, l31, fma(-l30, l30, a33)));
    const double r3 = pivot_rsqrt(a33, clean);
    all_clean = all_clean && clean;
    // M = L^{-1} (lower triangular)
    const double m00 = r0, m11 = r1, m22 = r2, m33 = r3;
    const double m10 = -(l10 * m00) * r1;
    const double m21 = -(l21 * m11) * r2;
    const double m32 = -(l32 * m22) * r3;
    const double m20 = -fma(l21, m10, l20 * m00) * r2;
    const double m31 = -fma(l32, m21, l31 * m11) * r3;
    const double m30 = -fma(l32, m20, fma(l31, m10, l30 * m00)) * r3;
    double y0 = 0.0, y1 = 0.0, y2 = 0.0, y3 = 0.0;     // block J of the forward substitution
    if (rhs) {
      double b0, b1, b2, b3;
      ld4(rhs + (J << 2), b0, b1, b2, b3);
      y0 = m00 * b0; y1 = fma(m11, b1, m10 * b0); y2 = fma(m22, b2, fma(m21, b1, m20 * b0));
      y3 = fma(m31, b1, m30 * b0) + fma(m33, b3, m32 * b2);
    }
    __syncwarp(gm);                                  // every lane has read the block before it is overwritten
    JMPC_TOCK(tc_, 13);
    // ---- panel below the block: X = A L^{-T}, one matrix row per lane
    const int prow = (nb - J - 1) << 2;
    for (int r = gl; r < prow; r += G) {
      double* row = K + tile_off(J + 1 + (r >> 2), J) + ((r & 3) << 2);
      double a0, a1, a2, a3;
      ld4(row, a0, a1, a2, a3);
      const double x0 = a0 * m00, x1 = fma(a1, m11, a0 * m10), x2 = fma(a2, m22, fma(a1, m21, a0 * m20));
      const double x3 = fma(a3, m33, fma(a2, m32, fma(a1, m31, a0 * m30)));
      st4(row, x0, x1, x2, x3);
      if (rhs) {
        double* bi = rhs + ((J + 1) << 2) + r;
        *bi -= fma(x1, y1, x0 * y0) + fma(x3, y3, x2 * y2);
      }
    }
    if (gl == 0) {
      // the diagonal tile receives the inverse of the factor's diagonal block: nothing reads the block itself again
      // (the panel, the trailing update and the triangular sweeps all work with the inverse)
      double* Mw = K + tile_off(J, J);
      st4(Mw, m00, 0.0, 0.0, 0.0); st4(Mw + 4, m10, m11, 0.0, 0.0);
      st4(Mw + 8, m20, m21, m22, 0.0); st4(Mw + 12, m30, m31, m32, m33);
      if (rhs) st4(rhs + (J << 2), y0, y1, y2, y3);
    }
    __syncwarp(gm);
    JMPC_TOCK(tc_, 14);
    // ---- trailing update: C(I, Kc) -= L(I, J) L(Kc, J)'; a task is half a tile (two rows), so the 45 / 36 / 28 ...
    // tiles of the first block columns fill the 32 lanes better than whole tiles would
    const int m = nb - J - 1, ntasks = m * (m + 1);
    const chol_task* tasks = lut + chol_lut_offset(m);
    for (int q = gl; q < ntasks; q += G) {
      const unsigned e = tasks[q];
      const unsigned sLI = e & 255u, sLK = (e >> 8) & 255u, sC = sLI + ((e >> 16) & 15u), hb = (e >> 16) & 0x40u;
      const char* Kh = reinterpret_cast<const char*>(K) + hb;          // rows 2h, 2h + 1 of a tile start 64 h bytes in
      const double* LI = reinterpret_cast<const double*>(Kh + sLI * (unsigned)(kTS * sizeof(double)));
      const double* LK = K + sLK * (unsigned)kTS;
      double* C = reinterpret_cast<double*>(const_cast<char*>(Kh) + sC * (unsigned)(kTS * sizeof(double)));
      double k00, k01, k02, k03, k10, k11, k12, k13, k20, k21, k22, k23, k30, k31, k32, k33;
      ld4(LK, k00, k01, k02, k03); ld4(LK + 4, k10, k11, k12, k13);
      ld4(LK + 8, k20, k21, k22, k23); ld4(LK + 12, k30, k31, k32, k33);
#pragma unroll
      for (int r = 0; r < 2; ++r) {
        double x0, x1, x2, x3, c0, c1, c2, c3;
        ld4(LI + 4 * r, x0, x1, x2, x3);
        ld4(C + 4 * r, c0, c1, c2, c3);
        c0 = fma(-x3, k03, fma(-x2, k02, fma(-x1, k01, fma(-x0, k00, c0))));
        c1 = fma(-x3, k13, fma(-x2, k12, fma(-x1, k11, fma(-x0, k10, c1))));
        c2 = fma(-x3, k23, fma(-x2, k22, fma(-x1, k21, fma(-x0, k20, c2))));
        c3 = fma(-x3, k33, fma(-x2, k32, fma(-x1, k31, fma(-x0, k30, c3))));
        st4(C + 4 * r, c0, c1, c2, c3);
      }
    }
    __syncwarp(gm);
    JMPC_TOCK(tc_, 15);
  }
  return all_clean;
}

// Solve L L' x = b in place; b has 4*nb entries in shared memory (16-byte aligned).  The two sweeps are separate
// functions: the forward one can also ride along with the factorisation (chol_tiles).
template <int G>
__device__ inline void solve_forward_tiles(const double* K, double* b, int nb, int gl, unsigned gm) {
  const int n4 = nb << 2;
  JMPC_PRAGMA_SWEEP
  for (int J = 0; J < nb; ++J) {                      // forward: L y = b
    const double* Mw = K + tile_off(J, J);
    double b0, b1, b2, b3;
    ld4(b + (J << 2), b0, b1, b2, b3);
    const double m00 = Mw[0];
    const double2 mr1 = *reinterpret_cast<const double2*>(Mw + 4);
    const double m10 = mr1.x, m11 = mr1.y;
    double m20, m21, m22, mpad, m30, m31, m32, m33;
    ld4(Mw + 8, m20, m21, m22, mpad); ld4(Mw + 12, m30, m31, m32, m33);
    (void)mpad;
    const double y0 = m00 * b0, y1 = fma(m11, b1, m10 * b0), y2 = fma(m22, b2, fma(m21, b1, m20 * b0));
    const double y3 = fma(m33, b3, fma(m32, b2, fma(m31, b1, m30 * b0)));
    __syncwarp(gm);
    for (int i = ((J + 1) << 2) + gl; i < n4; i += G) {
      double l0, l1, l2, l3;
      ld4(K + tile_off(i >> 2, J) + ((i & 3) << 2), l0, l1, l2, l3);
      b[i] = fma(-l3, y3, fma(-l2, y2, fma(-l1, y1, fma(-l0, y0, b[i]))));
    }
    if (gl == 0) st4(b + (J << 2), y0, y1, y2, y3);
    __syncwarp(gm);
  }
}
template <int G>
__device__ inline void solve_backward_tiles(const double* K, double* b, int nb, int gl, unsigned gm) {
  JMPC_PRAGMA_SWEEP
  for (int J = nb - 1; J >= 0; --J) {                 // backward: L' x = y
    const double* Mw = K + tile_off(J, J);
    double y0, y1, y2, y3;
    ld4(b + (J << 2), y0, y1, y2, y3);
    const double m00 = Mw[0];
    const double2 mr1 = *reinterpret_cast<const double2*>(Mw + 4);
    const double m10 = mr1.x, m11 = mr1.y;
    double m20, m21, m22, mpad, m30, m31, m32, m33;
    ld4(Mw + 8, m20, m21, m22, mpad); ld4(Mw + 12, m30, m31, m32, m33);
    (void)mpad;
    const double x3 = m33 * y3, x2 = fma(m32, y3, m22 * y2), x1 = fma(m31, y3, fma(m21, y2, m11 * y1));
    const double x0 = fma(m30, y3, fma(m20, y2, fma(m10, y1, m00 * y0)));
    __syncwarp(gm);
    for (int i = gl; i < (J << 2); i += G) {
      const double* col = K + tile_off(J, i >> 2) + (i & 3);
      b[i] = fma(-col[12], x3, fma(-col[8], x2, fma(-col[4], x1, fma(-col[0], x0, b[i]))));
    }
    if (gl == 0) st4(b + (J << 2), x0, x1, x2, x3);
    __syncwarp(gm);
  }
}

// ---- triangular sweeps with the vector in registers (the low-latency kernels) ------------------------------------
// Lane gl of the group owns entries gl (lo) and gl + G (hi) of the right-hand side (4 nb <= 2 G), for the whole sweep.
// Block J's four entries reach every lane by shuffle from their owners, every lane multiplies them by the inverted
// diagonal block redundantly and updates its own rows from the factor (read-only here), and the owners replace their
// entries by the block's solution.  Nothing goes through shared memory but the factor, so a sweep needs no
// __syncwarp at all; the shared-memory version above pays, per block column, a store -> synchronise -> load round
// trip for the block, a load / store of every row's entry and two synchronisations.  Measured (T = 20): on a warp that
// runs alone the three sweeps of an interior-point iteration drop from 14.3 k to 8.7 k cycles (iteration 46.2 k ->
// 41.5 k); on a full batch the kernel gets 3-6 % slower (more shuffles and redundant loads, and with four warps per
// scheduler one warp's dependency chains are not the limit).  So these sweeps serve the launches that leave the SMs
// nearly empty (a single ego's step) and the shared-memory ones everything else; both produce the same bits (the
// arithmetic per entry is the same sequence of FMAs, checked by the self-test kernel).
// One block column is branch-free: a lane without a row to update loads from the block's own diagonal tile instead
// (always a valid address) and its result is dropped by a select, the owners pick their component by selects.  The
// warp issues in order, so a divergent `if` around the loads is not just a branch: it ends the basic block, and ptxas
// can then no longer start the factor's loads in the shadow of the shuffles (750 instead of 290 cycles per block
// column with branches).
template <int G>
__device__ __forceinline__ void sweep_load(const double* b, int n4, int gl, double& lo, double& hi) {
  lo = (gl < n4) ? b[gl] : 0.0;
  hi = (gl + G < n4) ? b[gl + G] : 0.0;
}
template <int G>
__device__ __forceinline__ void sweep_store(double* b, int n4, int gl, double lo, double hi) {
  if (gl < n4) b[gl] = lo;
  if (gl + G < n4) b[gl + G] = hi;
}
// the inverse of L's diagonal block J (lower triangular, row major in the diagonal tile)
struct DiagInv { double m00, m10, m11, m20, m21, m22, m30, m31, m32, m33; };
__device__ __forceinline__ DiagInv load_diag_inv(const double* Mw) {
  DiagInv d;
  d.m00 = Mw[0];
  const double2 r1 = *reinterpret_cast<const double2*>(Mw + 4);
  d.m10 = r1.x; d.m11 = r1.y;
  double pad;
  ld4(Mw + 8, d.m20, d.m21, d.m22, pad); ld4(Mw + 12, d.m30, d.m31, d.m32, d.m33);
  (void)pad;
  return d;
}
template <int G, bool HI>
__device__ __forceinline__ void forward_block(const double* K, int J, int n4, int gl, unsigned gm, const double* row_lo,
                                              const double* row_hi, double& lo, double& hi) {
  double& own = HI ? hi : lo;                      // the register that holds block J's entries on their owners
  const int l0 = (J << 2) & (G - 1);
  const double* Mw = K + tile_off(J, J);
  const int first = (J + 1) << 2;                  // rows below the block
  const bool act_lo = !HI && gl >= first && gl < n4, act_hi = gl + G >= first && gl + G < n4;
  double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0, c0, c1, c2, c3;
  if (!HI) ld4(act_lo ? row_lo + J * kTS : Mw, a0, a1, a2, a3);
  ld4(act_hi ? row_hi + J * kTS : Mw, c0, c1, c2, c3);
  const DiagInv d = load_diag_inv(Mw);
  const double e0 = __shfl_sync(gm, own, l0, G), e1 = __shfl_sync(gm, own, l0 + 1, G);
  const double e2 = __shfl_sync(gm, own, l0 + 2, G), e3 = __shfl_sync(gm, own, l0 + 3, G);
  const double y0 = d.m00 * e0, y1 = fma(d.m11, e1, d.m10 * e0), y2 = fma(d.m22, e2, fma(d.m21, e1, d.m20 * e0));
  const double y3 = fma(d.m33, e3, fma(d.m32, e2, fma(d.m31, e1, d.m30 * e0)));
  if (!HI) {
    const double v = fma(-a3, y3, fma(-a2, y2, fma(-a1, y1, fma(-a0, y0, lo))));
    lo = act_lo ? v : lo;
  }
  {
    const double v = fma(-c3, y3, fma(-c2, y2, fma(-c1, y1, fma(-c0, y0, hi))));
    hi = act_hi ? v : hi;
  }
  const double s01 = (gl & 1) ? y1 : y0, s23 = (gl & 1) ? y3 : y2, mine = (gl & 2) ? s23 : s01;
  own = ((gl >> 2) == (l0 >> 2)) ? mine : own;
}
// L y = b
template <int G>
__device__ __forceinline__ void solve_forward_regs(const double* K, int nb, int gl, unsigned gm, double& lo, double& hi) {
  const int n4 = nb << 2;
  const double* row_lo = K + tile_off(gl >> 2, 0) + ((gl & 3) << 2);
  const double* row_hi = K + tile_off((gl + G) >> 2, 0) + ((gl & 3) << 2);
  const int nlo = nb < G / 4 ? nb : G / 4;
#pragma unroll 1
  for (int J = 0; J < nlo; ++J) forward_block<G, false>(K, J, n4, gl, gm, row_lo, row_hi, lo, hi);
#pragma unroll 1
  for (int J = G / 4; J < nb; ++J) forward_block<G, true>(K, J, n4, gl, gm, row_lo, row_hi, lo, hi);
}
template <int G, bool HI>
__device__ __forceinline__ void backward_block(const double* K, int J, int n4, int gl, unsigned gm, int col_lo, int col_hi,
                                               double& lo, double& hi) {
  double& own = HI ? hi : lo;
  const int l0 = (J << 2) & (G - 1);
  const double* Krow = K + tile_off(J, 0);          // block row J of the factor
  const double* Mw = Krow + J * kTS;
  const int below = J << 2;                          // entries above the block: i < 4 J
  const bool act_lo = gl < below && gl < n4, act_hi = HI && gl + G < below;
  const double* pa = act_lo ? Krow + col_lo : Mw;
  const double a0 = pa[0], a1 = pa[4], a2 = pa[8], a3 = pa[12];
  double c0 = 0.0, c1 = 0.0, c2 = 0.0, c3 = 0.0;
  if (HI) { const double* pc = act_hi ? Krow + col_hi : Mw; c0 = pc[0]; c1 = pc[4]; c2 = pc[8]; c3 = pc[12]; }
  const DiagInv d = load_diag_inv(Mw);
  const double y0 = __shfl_sync(gm, own, l0, G), y1 = __shfl_sync(gm, own, l0 + 1, G);
  const double y2 = __shfl_sync(gm, own, l0 + 2, G), y3 = __shfl_sync(gm, own, l0 + 3, G);
  const double x3 = d.m33 * y3, x2 = fma(d.m32, y3, d.m22 * y2), x1 = fma(d.m31, y3, fma(d.m21, y2, d.m11 * y1));
  const double x0 = fma(d.m30, y3, fma(d.m20, y2, fma(d.m10, y1, d.m00 * y0)));
  {
    const double v = fma(-a3, x3, fma(-a2, x2, fma(-a1, x1, fma(-a0, x0, lo))));
    lo = act_lo ? v : lo;
  }
  if (HI) {
    const double v = fma(-c3, x3, fma(-c2, x2, fma(-c1, x1, fma(-c0, x0, hi))));
    hi = act_hi ? v : hi;
  }
  const double s01 = (gl & 1) ? x1 : x0, s23 = (gl & 1) ? x3 : x2, mine = (gl & 2) ? s23 : s01;
  own = ((gl >> 2) == (l0 >> 2)) ? mine : own;
}
// L' x = y
template <int G>
__device__ __forceinline__ void solve_backward_regs(const double* K, int nb, int gl, unsigned gm, double& lo, double& hi) {
  const int n4 = nb << 2;
  const int col_lo = (gl >> 2) * kTS + (gl & 3), col_hi = ((gl + G) >> 2) * kTS + (gl & 3);
#pragma unroll 1
  for (int J = nb - 1; J >= G / 4; --J) backward_block<G, true>(K, J, n4, gl, gm, col_lo, col_hi, lo, hi);
  const int nlo = nb < G / 4 ? nb : G / 4;
#pragma unroll 1
  for (int J = nlo - 1; J >= 0; --J) backward_block<G, false>(K, J, n4, gl, gm, col_lo, col_hi, lo, hi);
}

// y = P x for a symmetric P on tiles (diagonal tiles stored full); lane owns rows lane and lane + 32.
__device__ inline void symv_tiles(const double* P, const double* x, int nb, int lane,
                                  double& y0, double& y1) {
  const int n4 = nb << 2;
  y0 = 0.0; y1 = 0.0;
#pragma unroll
  for (int pass = 0; pass < 2; ++pass) {
    const int i = lane + 32 * pass;
    if (i < n4) {
      const int I = i >> 2, r = i & 3;
      double acc0 = 0.0, acc1 = 0.0;
      for (int J = 0; J <= I; ++J) {
        double p0, p1, p2, p3, x0, x1, x2, x3;
        ld4(P + tile_off(I, J) + (r << 2), p0, p1, p2, p3);
        ld4(x + (J << 2), x0, x1, x2, x3);
        acc0 = fma(p1, x1, fma(p0, x0, acc0));
        acc1 = fma(p3, x3, fma(p2, x2, acc1));
      }
      for (int J = I + 1; J < nb; ++J) {
        const double* col = P + tile_off(J, I) + r;
        double x0, x1, x2, x3;
        ld4(x + (J << 2), x0, x1, x2, x3);
        acc0 = fma(col[4], x1, fma(col[0], x0, acc0));
        acc1 = fma(col[12], x3, fma(col[8], x2, acc1));
      }
      if (pass == 0) y0 = acc0 + acc1; else y1 = acc0 + acc1;
    }
  }
}


// ---- rows k and T + k per lane ------------------------------------------------------------------------------------
// The stage rows keep their part of an n = 2T vector in registers: lane k < T owns entries k (the acceleration part)
// and T + k (the steering part).  The symmetric matvec below produces its result in that layout.  (The
// triangular sweeps with the vector in registers further up serve the low-latency kernels only.)
template <int G>
__device__ inline void solve_tiles(const double* K, double* b, int nb, int gl, unsigned gm) {
  solve_forward_tiles<G>(K, b, nb, gl, gm);
  solve_backward_tiles<G>(K, b, nb, gl, gm);
}

// y = P x for a symmetric P on tiles (diagonal tiles stored full), x in shared memory (4 * nb entries); lane k < T gets
// rows k (y0) and T + k (y1), four independent accumulators per row.
template <int NB>
__device__ __forceinline__ void symv_rows(const double* P, const double* x, int T, int nb_rt, int gl, double& y0,
                                          double& y1) {
  const int nb = (NB > 0) ? NB : nb_rt;
  const bool own = gl < T;
  const int ia = own ? gl : 0, ib = own ? T + gl : 0;
  const int Ia = ia >> 2, Ib = ib >> 2;
  const double* rowa = P + tile_off(Ia, 0) + ((ia & 3) << 2);
  const double* rowb = P + tile_off(Ib, 0) + ((ib & 3) << 2);
  const double* cola = P + Ia * kTS + (ia & 3);
  const double* colb = P + Ib * kTS + (ib & 3);
  double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0, b0 = 0.0, b1 = 0.0, b2 = 0.0, b3 = 0.0;
  JMPC_PRAGMA_SYMV
  for (int J = 0; J < nb; ++J) {
    double x0, x1, x2, x3, p0, p1, p2, p3;
    ld4(x + (J << 2), x0, x1, x2, x3);
    const int tj = ((J * (J + 1)) >> 1) * kTS;
    if (J <= Ia) ld4(rowa + J * kTS, p0, p1, p2, p3);
    else { const double* col = cola + tj; p0 = col[0]; p1 = col[4]; p2 = col[8]; p3 = col[12]; }
    a0 = fma(p0, x0, a0); a1 = fma(p1, x1, a1); a2 = fma(p2, x2, a2); a3 = fma(p3, x3, a3);
    if (J <= Ib) ld4(rowb + J * kTS, p0, p1, p2, p3);
    else { const double* col = colb + tj; p0 = col[0]; p1 = col[4]; p2 = col[8]; p3 = col[12]; }
    b0 = fma(p0, x0, b0); b1 = fma(p1, x1, b1); b2 = fma(p2, x2, b2); b3 = fma(p3, x3, b3);
  }
  y0 = own ? (a0 + a1) + (a2 + a3) : 0.0;
  y1 = own ? (b0 + b1) + (b2 + b3) : 0.0;
}

// Self-test kernel: one lane group packs a dense symmetric matrix into tiles, multiplies, factors and solves.  With
// G < 32 every group of the warp works on its own copy in its own shared-memory region, and group `which` reports.
template <int G>
__global__ void linalg_selftest_kernel(int n, int which, const double* __restrict__ A, const double* __restrict__ b,
                                       const double* __restrict__ x, double* __restrict__ sol,
                                       double* __restrict__ prod, int* __restrict__ ok) {
  extern __shared__ double sm[];
  const int lane = threadIdx.x & 31, gl = lane & (G - 1), sub = lane / G, nb = nblk(n), n4 = nb << 2;
  const unsigned gm = (G == 32) ? kFullMask : (((1u << (G & 31)) - 1u) << (sub * G));
  const int per_group = tiles_doubles(n) + 2 * n4;
  double* K = sm + sub * per_group;
  double* rhs = K + tiles_doubles(n);
  double* xv = rhs + n4;
  chol_task* lut = reinterpret_cast<chol_task*>(sm + (32 / G) * per_group);
  chol_lut_build(lut, nb, lane, 32);
  __syncwarp();
  const bool report = sub == which;
  auto pack = [&]() {
    for (int e = gl; e < tiles_doubles(n); e += G) K[e] = 0.0;
    __syncwarp(gm);
    for (int e = gl; e < n4 * n4; e += G) {
      const int i = e / n4, j = e % n4;
      if (j > i && (i >> 2) != (j >> 2)) continue;
      K[elem_off(i, j)] = (i < n && j < n) ? A[i * n + j] : ((i == j) ? 1.0 : 0.0);
    }
    __syncwarp(gm);
  };
  pack();
  for (int i = gl; i < n4; i += G) { rhs[i] = (i < n) ? b[i] : 0.0; xv[i] = (i < n) ? x[i] : 0.0; }
  __syncwarp(gm);
  if (G == 32) {                                    // the plain tiled matvec only exists for a whole warp
    double y0, y1;
    symv_tiles(K, xv, nb, lane, y0, y1);
    if (lane < n) prod[lane] = y0;
    if (lane + 32 < n) prod[lane + 32] = y1;
    __syncwarp(gm);
  }
  const bool good = chol_tiles<G>(K, nb, gl, gm, lut);
  {
    // the solve through shared memory; then the same solve with the vector in registers must give the same bits
    // (NaN is reported otherwise)
    double lo, hi, lo2, hi2;
    sweep_load<G>(rhs, n4, gl, lo2, hi2);
    solve_tiles<G>(K, rhs, nb, gl, gm);
    __syncwarp(gm);
    sweep_load<G>(rhs, n4, gl, lo, hi);
    solve_forward_regs<G>(K, nb, gl, gm, lo2, hi2);
    solve_backward_regs<G>(K, nb, gl, gm, lo2, hi2);
    const bool same = (lo == lo2 || (lo != lo && lo2 != lo2)) && (hi == hi2 || (hi != hi && hi2 != hi2));
    __syncwarp(gm);
    sweep_store<G>(rhs, n4, gl, same ? lo : nan(""), same ? hi : nan(""));
  }
  __syncwarp(gm);
  if (report) {
    for (int i = gl; i < n; i += G) sol[i] = rhs[i];
    if (gl == 0) *ok = good ? 1 : 0;
  }
  // the matvec in the solver's layout (even n, n / 2 <= G: lane k owns rows k and n/2 + k); its result goes behind
  // the other one: prod[n .. 2n)
  if ((n & 1) == 0 && (n >> 1) <= G) {
    const int T = n >> 1;
    __syncwarp(gm);
    pack();                                         // symv needs the matrix again
    double q0, q1;
    symv_rows<0>(K, xv, T, nb, gl, q0, q1);
    if (report && gl < T) { prod[n + gl] = q0; prod[n + T + gl] = q1; }
  }
}

}  // namespace jmpc
