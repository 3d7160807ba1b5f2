// Thin SVD (np.linalg.svd(full_matrices=False), reference call site
// tensor.py:915) for float64 / complex128, entirely on the device.
//
//   A (m x n)  ->  U (m x k), S (k, descending), Vh (k x n),  k = min(m, n)
//
// Algorithm (no Gram-matrix shortcut on A itself -- that would square the
// condition number and lose the small singular values):
//   1. QR pre-reduction to a k x k triangular factor (qr.cu): A = Q R when
//      m >= n, A^H = Q R otherwise.
//   2. One-sided (Hestenes) block Jacobi on X = R^H:  X V = Y with mutually
//      orthogonal columns, so R^H = Y V^H.  Working on R^H rather than R is the
//      Drmac-Veselic preconditioning: its columns are much closer to
//      orthogonal and the sweep count drops.  (Projection mode of a tall or
//      square matrix needs the LEFT vectors of R: it factors R^H = Q2 R2 once
//      more and runs on X = R2^H -- see svd_impl.)
//   3. sigma_j = |Y[:, j]|, sorted on the device; U / Vh assembled with one
//      DMMA GEMM (Q V) and one gather/scale kernel (Y / sigma).
//
// Jacobi layout: the "columns" being orthogonalised are stored as ROWS of
// Xt (k x k) and Vt (k x k, starts as I), so a column is contiguous.  Columns
// are grouped in blocks of JB = 16.  One sweep is a round-robin tournament
// over block pairs (nblk - 1 rounds, nblk/2 independent pairs per round); one
// kernel launch per round -- the first round of a sweep also rotates the column
// pairs INSIDE its blocks (15 steps before its 16 cross steps, on the same Gram
// matrix), so every column pair is rotated exactly once per sweep.  A pair is
// handled by one thread-block CLUSTER:
//   * the 32 rows of the pair (both Xt and Vt, viewed as one long row) are cut
//     into chunks of CH elements; each CTA of the cluster owns chunks
//     crank, crank+S, ... and keeps the last Xt chunk it read resident in
//     shared memory;
//   * each CTA forms the partial 32 x 32 Gram matrix of its Xt chunks -- Hermitian:
//     only the 10 upper 8 x 8 tiles, complex products in the 3-multiplication
//     form (3 DMMAs instead of 4) -- and the partials are summed through
//     distributed shared memory (two cluster barriers, identical summation
//     order everywhere, so every CTA holds the same bits);
//   * every CTA runs the same parallel-ordered two-sided Jacobi steps on the
//     Gram matrix in shared memory (16 or 15 steps of 16 disjoint rotations):
//     one warp builds the rotations of step s + 1 while five warps update the
//     136 upper 2 x 2 blocks of G for step s and two warps carry this CTA's
//     share of the rows of the accumulated 32 x 32 unitary W (the rows are
//     split over the cluster and exchanged once, in the last step); the
//     rotation angles are exactly those of one-sided Jacobi on the columns;
//   * every CTA applies W to its chunks (rows_new = W^T rows, DMMA) and stores
//     them by TMA.
// Convergence: every rotation records whether its cosine |x_p^H x_q| / (|x_p||x_q|)
// exceeded tol and whether it exceeded 1e-7 (atomicOr on a device word); a
// one-thread kernel closes each sweep and sets a device flag that turns the
// remaining queued launches into no-ops, so the host only synchronises every
// few sweeps.  tol = sqrt(k) * eps (as LAPACK's
// xGESVJ).  HBM/L2 traffic per round: Xt and Vt read once and written once.
#include <cooperative_groups.h>
#include <math.h>
#include <stdlib.h>

#include <mutex>
#include <vector>

#include "common.cuh"

namespace cg = cooperative_groups;

namespace tnb {

int gemm(int dtype, int opA, int opB, int64_t M, int64_t N, int64_t K, double ar, double ai, const void* A,
         int64_t lda, int64_t sA, const void* B, int64_t ldb, int64_t sB, double br, double bi, void* C, int64_t ldc,
         int64_t sC, int64_t batch, cudaStream_t st);
int qr(int dtype, int64_t m, int64_t n, const void* A, int64_t lda, void* Q, void* R, void* ws, int scale_mode,
       double** scale_out, cudaStream_t st);
size_t qr_workspace(int dtype, int64_t m, int64_t n);
int permute_view(int dtype, const void* in, int rank, const int64_t* oshape, const int64_t* istride, void* out,
                 double ar, double ai, int conj, cudaStream_t st);

constexpr int JB = 16;       // Jacobi block width (columns)
constexpr int JP = 2 * JB;   // columns in a block pair
constexpr int JT = 512;      // threads per CTA: 2 x (JP/2)^2 -- G blocks on the first half, W blocks on the second
constexpr int JGP = JP + 4;  // pitch of W / Gram partials in shared memory (= 4 mod 8: conflict-free DMMA fragments)
constexpr int JGG = JP + 1;  // pitch of G (rotation phase only): row AND column accesses are conflict-free
constexpr int JWP = JP + 2;  // pitch of W (= 2 mod 8): the apply reads it transposed (fragment [k = p][m = q]) without conflicts
constexpr int JSP = JP + 4;  // pitch of the re + im table of W (doubles; = 4 mod 32: the four k rows of a fragment hit different banks)
constexpr int JPAD = 4;      // panel pitch = CH + JPAD for the same reason
constexpr int MAX_SWEEPS = 60;

#ifdef TNB_EXP_STAMPS
__device__ long long g_jac_dbg[128];  // kernel experiments: clock64 stamps of the rotation phase (CTA 0)
#define JSTAMP(k, cond) do { if (blockIdx.x == 0 && (cond)) g_jac_dbg[k] = clock64(); } while (0)
#else
#define JSTAMP(k, cond) do { } while (0)
#endif
constexpr int GRAM_TILES = 10;  // upper 8 x 8 tiles of the 4 x 4 tile grid of a JP x JP Hermitian matrix
constexpr int GRAM_UPW = 5;     // Gram work units (tile, eighth of a chunk) per warp: 10 * 8 / 16

// t-th upper tile (row-major over gm <= gn) of the 4 x 4 tile grid
__device__ __forceinline__ void upper_tile(int t, int& gm, int& gn) {
  if (t < 4) { gm = 0; gn = t; }
  else if (t < 7) { gm = 1; gn = t - 3; }
  else if (t < 9) { gm = 2; gn = t - 5; }
  else { gm = 3; gn = 3; }
}

struct JacobiFlags {
  unsigned int state;  // rotation flags of the sweep in flight (see jacobi_finish_sweep_kernel)
  unsigned int pad0;
  int converged;
  int sweeps;  // completed sweeps (stops counting once converged)
  int bad;     // a non-finite singular value was produced (NaN / Inf input)
  int pad;
};

struct JacobiArgs {
  void* X;
  void* V;
  int64_t n;    // number of Jacobi columns (= rows of Xt, Vt, and row length of Vt)
  int64_t L;    // row length of Xt
  int64_t ldx, ldv;
  int64_t xstride, vstride;  // elements between consecutive matrices of a batch (blockIdx.y)
  int nblk;     // even number of column blocks (the last ones may be empty)
  int round;
  int diag;     // 1: rotate the pairs INSIDE each of the two blocks (15 steps); 0: the 16 x 16 cross pairs (16 steps); 2: both (31)
  int S;        // CTAs per cluster
  int CH;       // chunk length (power of two)
  int nx, nv;   // chunks per Xt row / per Vt row
  double tol;
  int fp32_rot; // 1: Gram-domain rotation phase in single precision where the pair's column norms allow it
  JacobiFlags* flags;
};

// round-robin tournament (circle method), n even: pair p of round r
__host__ __device__ __forceinline__ void rr_pair(int n, int r, int p, int& a, int& b) {
  if (p == 0) { a = n - 1; b = r; return; }
  a = (r + p) % (n - 1);
  b = (r - p + n - 1) % (n - 1);
}

template <typename T> __device__ __forceinline__ double abs_t(T v);
template <> __device__ __forceinline__ double abs_t<double>(double v) { return fabs(v); }
template <> __device__ __forceinline__ double abs_t<cplx>(cplx v) { return hypot(v.x, v.y); }

// Rotation J = [[c, sp], [-conj(sp), c]] acting on columns (x_p, x_q) -> (x_p, x_q) J that
// annihilates x_p^H x_q; identity when the pair is already orthogonal to `tol`.
template <typename T> struct Rot { typename Num<T>::real_t c; T sp; };

// Rotation J = [[c, sp], [-conj(sp), c]] for the pair (p, q) from alpha = G[p][p], beta = G[q][q],
// g = G[p][q].  With tau = (beta - alpha)/2 and rh = 1/sqrt(tau^2 + |g|^2) (cos 2theta = |tau| rh):
//   c^2 = (1 + |tau| rh)/2,  sp = g sign(tau) rh / (2 c)       (|theta| <= pi/4, the inner rotation)
// -- two rsqrt, no division, no square root (X has unit Frobenius norm: nothing overflows).
// The pair counts as converged when |g|^2 <= tol2 alpha beta; bit 0 of `state` records a rotation,
// bit 1 one whose squared cosine exceeded 1e-14 (see jacobi_finish_sweep_kernel).
//
// Single-precision form (T = float / float2, the FP32 rotation phase): the formula is invariant under a
// common scaling of (alpha, beta, g), so the three are first divided by max(alpha, beta) -- nothing
// under- or overflows whatever the size of the columns -- and the tests are made on the squared cosine
// itself.  MUFU approximations (rcp, rsqrt) are good enough: the rotation the MATRIX sees is
// re-normalised in double precision (load_rot), the Gram matrix only ever decides angles.
__device__ __forceinline__ float rcp_approx(float x) { float r; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
__device__ __forceinline__ float rsqrt_approx(float x) { float r; asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
__device__ __forceinline__ float2 sel(bool c, float2 a, float2 b) { return make_float2(c ? a.x : b.x, c ? a.y : b.y); }
__device__ __forceinline__ float sel(bool c, float a, float b) { return c ? a : b; }
__device__ __forceinline__ cplx sel(bool c, cplx a, cplx b) { return make_double2(c ? a.x : b.x, c ? a.y : b.y); }
__device__ __forceinline__ double sel(bool c, double a, double b) { return c ? a : b; }

template <typename T>
__device__ __forceinline__ Rot<T> make_rot_vals(typename Num<T>::real_t alpha, typename Num<T>::real_t beta, T gam,
                                                double tol2, unsigned& state) {
  typedef Num<T> N_;
  Rot<T> r;
  r.c = 1; r.sp = N_::zero();
  if constexpr (sizeof(typename N_::real_t) == 4) {
    // branch-free (this is the head of the dependent chain of a step): garbage from a zero column or an
    // already orthogonal pair (0 * inf, rsqrt(0)) is discarded by the final selects -- NaN compares false
    const float inv = rcp_approx(fmaxf(alpha, beta));
    const float a_ = alpha * inv, b_ = beta * inv;
    const T g_ = N_::scale(gam, inv);
    const float ag2 = N_::abs2(g_);
    const float cos2 = ag2 * rcp_approx(a_ * b_);
    const bool rot = alpha > 0.f && beta > 0.f && cos2 > (float)tol2;
    const float tau = 0.5f * (b_ - a_);
    const float rh = rsqrt_approx(fmaf(tau, tau, ag2));
    const float c2 = fmaf(0.5f * fabsf(tau), rh, 0.5f);
    const float rc = rsqrt_approx(c2);
    const T sp = N_::scale(g_, copysignf(0.5f * rh * rc, tau));
    r.c = rot ? c2 * rc : 1.f;
    r.sp = sel(rot, sp, N_::zero());
    state |= rot ? ((cos2 > 1e-14f) ? 3u : 1u) : 0u;
  } else {
    const double ag2 = N_::abs2(gam);
    const double ab = alpha * beta;
    if (ab > 0.0 && ag2 > tol2 * ab) {
      state |= (ag2 > 1e-14 * ab) ? 3u : 1u;
      const double tau = 0.5 * (beta - alpha);
      const double rh = rsqrt(fma(tau, tau, ag2));
      const double c2 = fma(0.5 * fabs(tau), rh, 0.5);
      const double rc = rsqrt(c2);
      r.c = c2 * rc;
      r.sp = N_::scale(gam, copysign(0.5 * rh * rc, tau));
    }
  }
  return r;
}
template <typename T>
__device__ __forceinline__ Rot<T> make_rot(const T* G, int p, int q, double tol2, unsigned& state) {
  typedef Num<T> N_;
  return make_rot_vals<T>(N_::real(G[p * JGG + p]), N_::real(G[q * JGG + q]), G[p * JGG + q], tol2, state);
}

// The rotation as the MATRIX sees it (W is accumulated in the precision of the data).  A rotation built
// in single precision has c^2 + |sp|^2 = 1 + d with |d| ~ 1e-7: scaling both by 1 - d/2 + 3 d^2/8
// (= (1 + d)^(-1/2) up to d^3) makes it unitary to double precision; the identity stays exact.
template <typename T, typename GT>
__device__ __forceinline__ Rot<T> load_rot(const typename Num<GT>::real_t* rc, const GT* rsp, int i) {
  Rot<T> r;
  if constexpr (sizeof(GT) == sizeof(T)) {
    r.c = rc[i]; r.sp = rsp[i];
  } else {
    const GT sp = rsp[i];
    double c = (double)rc[i];
    T s;
    if constexpr (sizeof(T) == 16) s = make_double2((double)sp.x, (double)sp.y); else s = (double)sp;
    const double d = fma(c, c, Num<T>::abs2(s)) - 1.0;
    const double k = fma(d, fma(d, 0.375, -0.5), 1.0);
    r.c = c * k;
    r.sp = Num<T>::scale(s, k);
  }
  return r;
}

// c * x + s * y  (c real, s complex or real), written as FMA chains
__device__ __forceinline__ double rot_mix(double c, double x, double s, double y) { return fma(c, x, s * y); }
__device__ __forceinline__ cplx rot_mix(double c, cplx x, cplx s, cplx y) {
  return make_double2(fma(c, x.x, fma(s.x, y.x, -(s.y * y.y))), fma(c, x.y, fma(s.x, y.y, s.y * y.x)));
}
__device__ __forceinline__ float rot_mix(float c, float x, float s, float y) { return fmaf(c, x, s * y); }
__device__ __forceinline__ float2 rot_mix(float c, float2 x, float2 s, float2 y) {
  return make_float2(fmaf(c, x.x, fmaf(s.x, y.x, -(s.y * y.y))), fmaf(c, x.y, fmaf(s.x, y.y, s.y * y.x)));
}

// Entry (r, k) of B' = J_a^H B J_b for the 2 x 2 block B = [[b00, b01], [b10, b11]]
// (J = [[c, sp], [-conj(sp), c]]); the same operation order as the full block update of the round kernel.
// r and k differ from lane to lane: written with selects, not branches.
template <typename T>
__device__ __forceinline__ T rotated_entry(T b00, T b01, T b10, T b11, const Rot<T>& Ra, const Rot<T>& Rb, int r, int k) {
  typedef Num<T> N_;
  const bool k1 = (k != 0), r1 = (r != 0);
  // column k of T1 = B J_b:  T1[:,0] = c B[:,0] - conj(sp) B[:,1],  T1[:,1] = c B[:,1] + sp B[:,0]
  const T sb = sel(k1, Rb.sp, N_::sub(N_::zero(), N_::conj(Rb.sp)));
  const T t0 = rot_mix(Rb.c, sel(k1, b01, b00), sb, sel(k1, b00, b01));
  const T t1 = rot_mix(Rb.c, sel(k1, b11, b10), sb, sel(k1, b10, b11));
  // row r of J_a^H T1,  J_a^H = [[c, -sp], [conj(sp), c]]
  const T sa = sel(r1, N_::conj(Ra.sp), N_::sub(N_::zero(), Ra.sp));
  return rot_mix(Ra.c, sel(r1, t1, t0), sa, sel(r1, t0, t1));
}

// ---- rotation phase of a round: parallel-ordered Jacobi rotations on G, accumulated in W ---------------------
// 16 (cross round) or 15 (diag round) steps of 16 disjoint rotations, one barrier per step.  Non-tensor
// FP64 instructions are the scarce resource here (measured: ~7 issue cycles per FP64 warp instruction on
// the busiest SM sub-partition decide the length of a step; warp w runs on sub-partition w & 3), hence:
//   warp 15 builds the rotations of step s + 1 WHILE the others apply those of step s: the three entries
//     of G^(s+1) a rotation needs follow from three 2 x 2 blocks of G^(s) and the two rotations of step s
//     that touch its columns (rotated_entry);
//   G' = J^H G J is Hermitian: warps 0, 4, 1, 5, 2 own the 136 blocks ta <= tb (thread = one 2 x 2 block,
//     B' = J_a^H B J_b) and store the mirror image too (G's pitch of 33 makes column stores conflict-free);
//   W' = W J acts on rows independently: with S = 2, 4 or 8 CTAs in the cluster each CTA only carries
//     32 / S rows of W through the steps (task = one row x the two columns of pair wb, dealt to up to ten
//     warps spread over the four sub-partitions); the last step writes its rows into every CTA of the
//     cluster (distributed shared memory), one cluster barrier follows.
// G and W ping-pong between two buffers, the rotations between rc/rsp[0] and [1].
// GT is the precision of the Gram domain (T itself, or its single-precision counterpart: see the call site).
struct RotCtx {
  const unsigned char (*rr)[JP / 2][2];
  const unsigned int (*nxt)[JP / 2];
  const unsigned char (*gt)[2];
  int nsteps, S, crank, w_tasks, w_first, g0, wi, nw;
  bool wsplit;
  double tol2;
};

template <typename T, typename GT>
__device__ __forceinline__ void rotation_phase(GT* G, int gbuf, T* W, typename Num<GT>::real_t (*s_rc)[JP / 2],
                                               GT (*s_rsp)[JP / 2], const RotCtx& cx, cg::cluster_group& cluster,
                                               unsigned& state, int warp, int lane) {
  typedef Num<T> N_;
  typedef Num<GT> NG;
  constexpr int ROTW = JT / 32 - 1;
  constexpr int NGB = JB * (JB + 1) / 2;   // 136 blocks
  const int nsteps = cx.nsteps;
  if (warp == ROTW && lane < JP / 2) {
    const Rot<GT> r = make_rot<GT>(G, cx.rr[0][lane][0], cx.rr[0][lane][1], cx.tol2, state);
    s_rc[0][lane] = r.c;
    s_rsp[0][lane] = r.sp;
  }
  // the rotation warp fetches its table word one step ahead (it does not depend on the data)
  unsigned nx_ahead = (warp == ROTW && lane < JP / 2 && nsteps > 1) ? cx.nxt[0][lane] : 0u;
  // a G task keeps its block (ta, tb) through all steps
  int g_ta = 0, g_tb = 0;
  if (cx.g0 >= 0 && cx.g0 + lane < NGB) { g_ta = cx.gt[cx.g0 + lane][0]; g_tb = cx.gt[cx.g0 + lane][1]; }
  __syncthreads();
  JSTAMP(0, threadIdx.x == 0);
  for (int step = 0; step < nsteps; ++step) {
    const int cur = step & 1;
    JSTAMP(1 + 4 * step, threadIdx.x == ROTW * 32);
    const GT* Gc = G + cur * gbuf;
    GT* Gn = G + (cur ^ 1) * gbuf;
    const T* Wc = W + cur * (JP * JGP);
    T* Wn = W + (cur ^ 1) * (JP * JGP);
    if (warp == ROTW) {
      if (lane < JP / 2 && step + 1 < nsteps) {
        const unsigned nx_ = nx_ahead;
        if (step + 2 < nsteps) nx_ahead = cx.nxt[step + 1][lane];
        const int p1 = nx_ & 31, q1 = (nx_ >> 5) & 31, p2 = (nx_ >> 10) & 31, q2 = (nx_ >> 15) & 31;
        const int a1 = (nx_ >> 20) & 15, a2 = (nx_ >> 24) & 15, r1 = (nx_ >> 28) & 1, r2 = (nx_ >> 29) & 1;
        Rot<GT> R1, R2;
        R1.c = s_rc[cur][a1]; R1.sp = s_rsp[cur][a1];
        R2.c = s_rc[cur][a2]; R2.sp = s_rsp[cur][a2];
        const GT al = rotated_entry<GT>(Gc[p1 * JGG + p1], Gc[p1 * JGG + q1], Gc[q1 * JGG + p1], Gc[q1 * JGG + q1], R1, R1, r1, r1);
        const GT be = rotated_entry<GT>(Gc[p2 * JGG + p2], Gc[p2 * JGG + q2], Gc[q2 * JGG + p2], Gc[q2 * JGG + q2], R2, R2, r2, r2);
        const GT ga = rotated_entry<GT>(Gc[p1 * JGG + p2], Gc[p1 * JGG + q2], Gc[q1 * JGG + p2], Gc[q1 * JGG + q2], R1, R2, r1, r2);
        const Rot<GT> r = make_rot_vals<GT>(NG::real(al), NG::real(be), ga, cx.tol2, state);
        s_rc[cur ^ 1][lane] = r.c;
        s_rsp[cur ^ 1][lane] = r.sp;
      }
      JSTAMP(2 + 4 * step, lane == 0);
    } else if (cx.g0 >= 0) {
      const int t = cx.g0 + lane;
      JSTAMP(80, threadIdx.x == 0 && step == 5);
      if (t < NGB) {
        const int ta = g_ta, tb = g_tb;
        const int pa = cx.rr[step][ta][0], qa = cx.rr[step][ta][1];
        const int pb = cx.rr[step][tb][0], qb = cx.rr[step][tb][1];
        Rot<GT> Ra, Rb;
        Ra.c = s_rc[cur][ta]; Ra.sp = s_rsp[cur][ta];
        Rb.c = s_rc[cur][tb]; Rb.sp = s_rsp[cur][tb];
        const GT msb = NG::sub(NG::zero(), NG::conj(Rb.sp));  // -conj(sp_b)
        const GT b00 = Gc[pa * JGG + pb], b01 = Gc[pa * JGG + qb], b10 = Gc[qa * JGG + pb], b11 = Gc[qa * JGG + qb];
#ifdef TNB_EXP_STAMPS
        if (blockIdx.x == 0 && threadIdx.x == 0 && step == 5) { g_jac_dbg[81] = clock64() + (long long)(NG::real(b00) * 0 + NG::real(b01) * 0 + NG::real(b10) * 0 + NG::real(b11) * 0 + Ra.c * 0 + Rb.c * 0); }
#endif
        // T1 = B J_b:  T1[:,0] = c B[:,0] - conj(sp) B[:,1],  T1[:,1] = sp B[:,0] + c B[:,1]
        const GT t00 = rot_mix(Rb.c, b00, msb, b01), t01 = rot_mix(Rb.c, b01, Rb.sp, b00);
        const GT t10 = rot_mix(Rb.c, b10, msb, b11), t11 = rot_mix(Rb.c, b11, Rb.sp, b10);
        // B' = J_a^H T1,  J_a^H = [[c, -sp], [conj(sp), c]]
        const GT msa = NG::sub(NG::zero(), Ra.sp), csa = NG::conj(Ra.sp);
        GT n00 = rot_mix(Ra.c, t00, msa, t10), n01 = rot_mix(Ra.c, t01, msa, t11);
        GT n10 = rot_mix(Ra.c, t10, csa, t00), n11 = rot_mix(Ra.c, t11, csa, t01);
#ifdef TNB_EXP_STAMPS
        if (blockIdx.x == 0 && threadIdx.x == 0 && step == 5) { g_jac_dbg[82] = clock64() + (long long)(NG::real(n00) * 0 + NG::real(n01) * 0 + NG::real(n10) * 0 + NG::real(n11) * 0); }
#endif
        if (ta == tb) {
          // diagonal block: real diagonal, exact zero where a rotation was applied
          n00 = NG::from(NG::real(n00), 0);
          n11 = NG::from(NG::real(n11), 0);
          if (Ra.c != 1 || NG::abs2(Ra.sp) != 0) { n01 = NG::zero(); n10 = NG::zero(); }
        } else {
          Gn[pb * JGG + pa] = NG::conj(n00); Gn[qb * JGG + pa] = NG::conj(n01);
          Gn[pb * JGG + qa] = NG::conj(n10); Gn[qb * JGG + qa] = NG::conj(n11);
        }
        Gn[pa * JGG + pb] = n00; Gn[pa * JGG + qb] = n01; Gn[qa * JGG + pb] = n10; Gn[qa * JGG + qb] = n11;
      }
      JSTAMP(3 + 4 * step, threadIdx.x == 0);
    } else if (cx.wi >= 0) {
      for (int task = cx.wi * 32 + lane; task < cx.w_tasks; task += cx.nw * 32) {
        const int row = cx.w_first + (task >> 4), wb = task & 15;
        const int pb = cx.rr[step][wb][0], qb = cx.rr[step][wb][1];
        const Rot<T> Rb = load_rot<T, GT>(s_rc[cur], s_rsp[cur], wb);
        const T msb = N_::sub(N_::zero(), N_::conj(Rb.sp));
        const T w0 = Wc[row * JWP + pb], w1 = Wc[row * JWP + qb];
        const T v0 = rot_mix(Rb.c, w0, msb, w1), v1 = rot_mix(Rb.c, w1, Rb.sp, w0);
        Wn[row * JWP + pb] = v0;
        Wn[row * JWP + qb] = v1;
        if (cx.wsplit && step == nsteps - 1) {
          // last step: the final rows also go to the other CTAs of the cluster (they never touch these rows)
          for (int q = 0; q < cx.S; ++q) {
            if (q == cx.crank) continue;
            T* Wr = cluster.map_shared_rank(Wn, q);
            Wr[row * JWP + pb] = v0;
            Wr[row * JWP + qb] = v1;
          }
        }
      }
      JSTAMP(4 + 4 * step, lane == 0 && cx.wi == 0);
    }
    __syncthreads();
  }
  JSTAMP(1 + 4 * nsteps, threadIdx.x == 0);
  if (cx.wsplit) cluster.sync();  // every CTA's rows of the final W have arrived; nobody left before its rows were written
}


// Shared memory: P[JP][CH+JPAD] | G[2][JP][JGP] | W[2][JP][JGP] | Gpart[GRAM_TILES][64] (this CTA's Gram partial)
template <typename T>
__global__ void __launch_bounds__(JT) jacobi_round_kernel(JacobiArgs a) {
  typedef Num<T> N_;
  constexpr bool CPLX = (sizeof(T) == 16);
  JSTAMP(70, threadIdx.x == 0);
  cg::cluster_group cluster = cg::this_cluster();
  const int S = a.S;
  const int crank = (int)cluster.block_rank();
  const int pair = blockIdx.x / S;
  int bI, bJ;
  rr_pair(a.nblk, a.round, pair, bI, bJ);
  if (bI > bJ) { const int t = bI; bI = bJ; bJ = t; }

  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int CH = a.CH, pitch = CH + JPAD;
  T* P = reinterpret_cast<T*>(smem_raw);
  T* G = P + (size_t)JP * pitch;   // G and W are double-buffered through the sweep: [2][JP][JGP] each
  T* W = G + 2 * JP * JGP;
  T* Gpart = W + 2 * JP * JGP;   // [GRAM_TILES][64]: own buffer, so that W is initialised before the cluster meets and
                                 // nobody has to wait for the remote readers of its partial

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, gq = lane >> 2, tq = lane & 3;
  // s_rr[step][a] = (smaller, larger) panel column of rotation pair a in step `step`.
  //   cross round: column a of block I meets column (a + step) mod 16 of block J      (16 steps)
  //   diag  round: two independent 16-player tournaments, one inside each block       (15 steps)
  // The in-block steps (once per sweep) plus nblk - 1 rounds of cross steps rotate every column pair exactly once per sweep.
  __shared__ unsigned char s_rr[2 * JB - 1][JP / 2][2];
  // s_pos[step][i] = 2 * (rotation pair holding panel column i in that step) + (1 if i is the larger one)
  __shared__ unsigned char s_pos[2 * JB - 1][JP];
  // s_nxt[step][a]: for rotation pair a of step + 1, the two pairs of `step` its columns come from, their
  // columns, and on which side of its pair each column sits (packed, see below)
  __shared__ unsigned int s_nxt[2 * JB - 1][JP / 2];
  // s_gt[t] = (ta, tb), ta <= tb: the 136 blocks of the upper block triangle of G (16 x 16 blocks of 2 x 2)
  __shared__ unsigned char s_gt[JB * (JB + 1) / 2][2];
  if (tid < JB * JB) {
    const int x = tid >> 4, y = tid & 15;
    if (x <= y) {
      const int t = x * JB - (x * (x - 1)) / 2 + (y - x);
      s_gt[t][0] = (unsigned char)x;
      s_gt[t][1] = (unsigned char)y;
    }
  }
  // a.diag = 2: the in-block steps and then the cross steps in ONE round (31 steps on the same Gram matrix: the
  // first launch of a sweep -- its pairs are those of cross round 0 -- so a sweep is nblk - 1 launches, not nblk)
  const int nd = (a.diag != 0) ? JB - 1 : 0;
  const int nsteps = nd + ((a.diag != 1) ? JB : 0);
  if (tid < nsteps * (JP / 2)) {
    const int st_ = tid / (JP / 2), pr_ = tid % (JP / 2);
    int x, y;
    if (st_ < nd) {
      rr_pair(JB, st_, pr_ & (JB / 2 - 1), x, y);
      if (pr_ >= JB / 2) { x += JB; y += JB; }
    } else {
      x = pr_;
      y = JB + ((pr_ + st_ - nd) & (JB - 1));
    }
    const int lo_ = x < y ? x : y, hi_ = x < y ? y : x;
    s_rr[st_][pr_][0] = (unsigned char)lo_;
    s_rr[st_][pr_][1] = (unsigned char)hi_;
    s_pos[st_][lo_] = (unsigned char)(2 * pr_);
    s_pos[st_][hi_] = (unsigned char)(2 * pr_ + 1);
  }
  __syncthreads();
  if (tid < (nsteps - 1) * (JP / 2)) {
    const int st_ = tid / (JP / 2), pr_ = tid % (JP / 2);
    const unsigned p1 = s_pos[st_][s_rr[st_ + 1][pr_][0]], p2 = s_pos[st_][s_rr[st_ + 1][pr_][1]];
    // packed: columns (p1, q1) / (p2, q2) of the two source pairs (5 bits each), the pairs themselves (4 bits
    // each) and the sides (1 bit each) -- one shared-memory word per rotation and step
    const unsigned a1 = p1 >> 1, a2 = p2 >> 1;
    s_nxt[st_][pr_] = (unsigned)s_rr[st_][a1][0] | ((unsigned)s_rr[st_][a1][1] << 5) | ((unsigned)s_rr[st_][a2][0] << 10) |
                      ((unsigned)s_rr[st_][a2][1] << 15) | (a1 << 20) | (a2 << 24) | ((p1 & 1u) << 28) | ((p2 & 1u) << 29);
  }
  // Programmatic dependent launch: everything above (index tables, barrier set-up) ran while the previous
  // round was still finishing; from here on this grid reads what that round wrote.  The next launch in the
  // stream may be scheduled as soon as every CTA of this grid got here (its CTAs take over the SMs one by
  // one as ours exit, and wait at this same point).
  JSTAMP(71, threadIdx.x == 0);
  griddep_wait();
  griddep_launch_dependents();
  JSTAMP(72, threadIdx.x == 0);
  JacobiFlags* const flags = a.flags + blockIdx.y;   // one matrix of the batch per grid row
  if (flags->converged) return;  // uniform over all clusters of this matrix
  T* Xg = reinterpret_cast<T*>(a.X) + (int64_t)blockIdx.y * a.xstride;
  T* Vg = a.V ? reinterpret_cast<T*>(a.V) + (int64_t)blockIdx.y * a.vstride : nullptr;

  auto grow = [&](int i) -> int64_t {  // global row of panel row i, or -1
    const int64_t r = (i < JB) ? ((int64_t)bI * JB + i) : ((int64_t)bJ * JB + (i - JB));
    return (r < a.n) ? r : (int64_t)-1;
  };
  auto chunk_geom = [&](int g, T*& base, int64_t& ld, int64_t& c0, int& len) {
    if (g < a.nx) { base = Xg; ld = a.ldx; c0 = (int64_t)g * CH; const int64_t rem = a.L - c0; len = (int)(rem < CH ? rem : CH); }
    else { base = Vg; ld = a.ldv; c0 = (int64_t)(g - a.nx) * CH; const int64_t rem = a.n - c0; len = (int)(rem < CH ? rem : CH); }
  };
  // Full chunks whose rows are 16-byte aligned move by TMA bulk copies (one 1-D copy per panel row,
  // issued by the lanes of warp 0, completion on an mbarrier); anything else by plain loads.
  // Each row moves as two copies (first and second half of the chunk) completing on two barriers, so the Gram
  // matrix of the first half is formed while the second half is still in flight.
  __shared__ __align__(8) uint64_t s_bar[2];
  uint32_t bar_phase = 0;
  if (tid == 0) { mbar_init(&s_bar[0], 1); mbar_init(&s_bar[1], 1); fence_mbar_init(); }
  auto bulk_ok = [&](const T* base, int64_t ld, int len) -> bool {
    return len == CH && (CPLX || ((ld & 1) == 0 && (reinterpret_cast<uintptr_t>(base) & 15) == 0));
  };
  // returns true when the chunk is in flight as bulk copies (wait_half(0), wait_half(1) follow), false when it
  // was loaded with plain loads (a __syncthreads() makes it visible)
  auto load_chunk = [&](int g) -> bool {
    T* base; int64_t ld, c0; int len;
    chunk_geom(g, base, ld, c0, len);
    if (bulk_ok(base, ld, len)) {
      if (warp == 0) {
        const int64_t r = grow(lane);
        const unsigned valid = __ballot_sync(0xffffffffu, r >= 0);
        const uint32_t hb = (uint32_t)((CH / 2) * sizeof(T));
        if (lane == 0) {
          mbar_arrive_expect_tx(&s_bar[0], (uint32_t)__popc(valid) * hb);
          mbar_arrive_expect_tx(&s_bar[1], (uint32_t)__popc(valid) * hb);
        }
        __syncwarp();
        if (r >= 0) {
          bulk_g2s(P + lane * pitch, base + r * ld + c0, hb, &s_bar[0]);
          bulk_g2s(P + lane * pitch + CH / 2, base + r * ld + c0 + CH / 2, hb, &s_bar[1]);
        }
      } else {
        for (int i = 0; i < JP; ++i) {  // rows past the end of the matrix are zero columns
          if (grow(i) >= 0) continue;
          for (int c = tid - 32; c < CH; c += JT - 32) P[i * pitch + c] = N_::zero();
        }
      }
      return true;
    }
    for (int idx = tid; idx < JP * CH; idx += JT) {
      const int i = idx / CH, c = idx - i * CH;
      const int64_t r = grow(i);
      T v = N_::zero();
      if (r >= 0 && c < len) v = base[r * ld + c0 + c];
      P[i * pitch + c] = v;
    }
    return false;
  };
  auto wait_half = [&](int h) { mbar_wait(&s_bar[h], bar_phase); };

  // ---- partial Gram matrix of this CTA's Xt chunks on the FP64 tensor pipe (DMMA.8x8x4) ----------
  // G[p][q] = sum_c conj(P[p][c]) P[q][c] is Hermitian: only the 10 upper 8 x 8 tiles of the 4 x 4 tile
  // grid are formed.  Work unit = (tile, eighth of the chunk); the 80 units are dealt 5 per warp, so a
  // warp touches at most two tiles (accumulator sets A and B, kept across the chunks of this CTA).
  // Both fragments are rows of P with c (= k) contiguous, lane (gq, tq) holds element [row0 + gq][k0 + tq].
  // Complex products use the 3-multiplication form (3 DMMAs per tile and k-step instead of 4):
  //   P1 = ar br, P2 = ai bi, P3 = (ar + ai)(br - bi):  re = P1 + P2,  im = P1 - P2 - P3.
  constexpr int NACC = CPLX ? 6 : 2;
  double accA[NACC], accB[NACC];
#pragma unroll
  for (int r = 0; r < NACC; ++r) { accA[r] = 0.0; accB[r] = 0.0; }
  const int u0 = warp * GRAM_UPW;
  const int tA = u0 >> 3, sA0 = u0 & 7;
  const int nA = (8 - sA0 < GRAM_UPW) ? (8 - sA0) : GRAM_UPW;  // units of this warp in tile tA
  const int nB = GRAM_UPW - nA;                                // ... and in tile tA + 1 (from its slice 0)
  int gmA, gnA, gmB, gnB;
  upper_tile(tA, gmA, gnA);
  upper_tile(nB > 0 ? tA + 1 : tA, gmB, gnB);
  const int SL = CH >> 3;  // chunk elements per unit (CH >= 32: a multiple of the DMMA k = 4)
  auto gram_units = [&](double (&acc)[NACC], int gm, int gn, int kbeg, int klen) {
    const T* pa = P + (gm * 8 + gq) * pitch + tq + kbeg;
    const T* pb = P + (gn * 8 + gq) * pitch + tq + kbeg;
#ifdef TNB_EXP_SKIP_GRAM
    klen = 4;
#endif
#pragma unroll 4
    for (int k0 = 0; k0 < klen; k0 += 4) {
      const T av = pa[k0], bv = pb[k0];
      if constexpr (CPLX) {
        dmma884(acc[0], acc[1], av.x, bv.x);
        dmma884(acc[2], acc[3], av.y, bv.y);
        dmma884(acc[4], acc[5], av.x + av.y, bv.x - bv.y);
      } else {
        dmma884(acc[0], acc[1], av, bv);
      }
    }
  };
  int resident = -1;
  // this warp's units by half of the chunk (slices 0-3 | 4-7); tile B always starts at slice 0 and nB <= 4
  const int a_lo_end = (sA0 + nA < 4) ? sA0 + nA : 4;           // tile A, first half:  slices [sA0, a_lo_end)
  const int a_hi_beg = (sA0 > 4) ? sA0 : 4;                     // tile A, second half: slices [a_hi_beg, sA0 + nA)
  for (int gch = crank; gch < a.nx; gch += S) {
    __syncthreads();
    const bool bulk = load_chunk(gch);
    __syncthreads();
    resident = gch;
    if (bulk) wait_half(0);
    if (sA0 < a_lo_end) gram_units(accA, gmA, gnA, sA0 * SL, (a_lo_end - sA0) * SL);
    if (nB > 0) gram_units(accB, gmB, gnB, 0, nB * SL);
    if (bulk) { wait_half(1); bar_phase ^= 1; }
    if (a_hi_beg < sA0 + nA) gram_units(accA, gmA, gnA, a_hi_beg * SL, (sA0 + nA - a_hi_beg) * SL);
  }
  JSTAMP(73, threadIdx.x == 0);
  // per-warp partial tiles -> slots (in the space of G), summed per tile in a fixed order -> Gpart
  T* slots = G;  // [16 warps][2][64]
  {
    auto flush = [&](const double (&acc)[NACC], int j) {
      T* dst = slots + (warp * 2 + j) * 64 + gq * 8 + 2 * tq;
      if constexpr (CPLX) {
        dst[0] = make_double2(acc[0] + acc[2], acc[0] - acc[2] - acc[4]);
        dst[1] = make_double2(acc[1] + acc[3], acc[1] - acc[3] - acc[5]);
      } else {
        dst[0] = acc[0];
        dst[1] = acc[1];
      }
    };
    flush(accA, 0);
    if (nB > 0) flush(accB, 1);
  }
  __syncthreads();
  for (int idx = tid; idx < GRAM_TILES * 64; idx += JT) {
    const int t = idx >> 6, e = idx & 63;
    int gm, gn;
    upper_tile(t, gm, gn);
    T sum = N_::zero();
    for (int w = (8 * t) / GRAM_UPW; w <= (8 * t + 7) / GRAM_UPW; ++w)
      sum = N_::add(sum, slots[(w * 2 + ((((w * GRAM_UPW) >> 3) == t) ? 0 : 1)) * 64 + e]);
    Gpart[idx] = sum;
  }
  for (int idx = tid; idx < JP * JP; idx += JT) {
    const int i = idx / JP, j = idx - i * JP;
    W[i * JWP + j] = (i == j) ? N_::one() : N_::zero();
  }
  cluster.sync();
  // sum of the cluster's partials (same order in every CTA: identical bits everywhere), mirrored into
  // the lower triangle; the diagonal is real
  for (int idx = tid; idx < GRAM_TILES * 64; idx += JT) {
    const int t = idx >> 6, e = idx & 63;
    int gm, gn;
    upper_tile(t, gm, gn);
    const int i = gm * 8 + (e >> 3), j = gn * 8 + (e & 7);
    if (i > j) continue;
    T part[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) part[q] = (q < S) ? cluster.map_shared_rank(Gpart, q)[idx] : N_::zero();
    T sum = part[0];
#pragma unroll
    for (int q = 1; q < 8; ++q) sum = N_::add(sum, part[q]);
    if (i == j) {
      G[i * JGG + i] = N_::from(N_::real(sum), 0.0);
    } else {
      G[i * JGG + j] = sum;
      G[j * JGG + i] = N_::conj(sum);
    }
  }
  // No second cluster barrier: nobody overwrites its partial in this launch, and no CTA leaves before the barrier that
  // ends the rotation phase (S > 1 always exchanges the rows of W there), by which time every reader is done.
  JSTAMP(74, threadIdx.x == 0);
  __syncthreads();

  // ---- parallel-ordered Jacobi rotations on G, accumulated in W (rotation_phase above) -------------------
  // The Gram matrix only decides ANGLES; the matrix itself sees W, which is accumulated in double precision
  // from re-normalised rotations.  So the whole Gram-domain chain (rotation -> three entries of the next G ->
  // next rotation, and the update of G) runs in SINGLE precision whenever the column norms of the pair lie
  // within a factor 2^20 of each other (squared norms within 2^40: every squared cosine down to tol^2 is
  // then representable): FP32 instructions issue every cycle where an FP64 one takes 5-8, and MUFU rcp /
  // rsqrt replace the 60-cycle double-precision ones.  Pairs with a wider spread (graded / rank-deficient
  // matrices) keep the double-precision phase -- the decision is taken from the reduced Gram matrix, which
  // holds the same bits in every CTA of the cluster.
  RotCtx cx;
  cx.rr = s_rr; cx.nxt = s_nxt; cx.gt = s_gt;
  cx.nsteps = nsteps; cx.S = S; cx.crank = crank;
  cx.wsplit = (S > 1) && (JP % S == 0) && (S <= JP / 2);
  const int w_rows = cx.wsplit ? JP / S : JP;   // rows of W carried by this CTA
  cx.w_first = cx.wsplit ? crank * w_rows : 0;
  cx.w_tasks = w_rows * (JP / 2);
  cx.nw = (cx.w_tasks + 31) / 32 < 10 ? (cx.w_tasks + 31) / 32 : 10;
  cx.g0 = -1; cx.wi = -1;
  {
    const int g_warp[5] = {0, 4, 1, 5, 2};
    const int w_warp[10] = {8, 9, 10, 6, 12, 13, 14, 3, 7, 11};   // sub-partitions 0 1 2 2 0 1 2 3 3 3: the rotation warp (15, sub-partition 3) keeps its issue port to itself as long as possible
    // (a step is ISSUE-bound: ~1250 issue slots of the five G warps, ~1200 of the four W warps, ~300 of the rotation
    // warp over four sub-partitions ~ 690 cycles against 770 measured; other placements of the roles -- G on 0 0 1 1 0
    // with W on 2, W next to the rotation warp, G on 0 1 2 3 0 -- came out at 29.40 / 28.78 / 28.94 us per round
    // against 28.73)
#pragma unroll
    for (int i = 0; i < 5; ++i) if (warp == g_warp[i]) cx.g0 = 32 * i;
#pragma unroll
    for (int i = 0; i < 10; ++i) if (warp == w_warp[i] && i < cx.nw) cx.wi = i;
  }
  cx.tol2 = a.tol * a.tol;
  unsigned state = 0;
  __shared__ double s_rc[2][JP / 2];
  __shared__ T s_rsp[2][JP / 2];
#ifndef TNB_EXP_SKIP_EIGEN
  bool use32 = false;
  double gscale = 1.0;
  if (a.fp32_rot) {
    const double d = N_::real(G[lane * JGG + lane]);   // JP == 32 == warp size
    double dmax = d, dmin = d > 0.0 ? d : 1e300;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      dmax = fmax(dmax, __shfl_xor_sync(0xffffffffu, dmax, o));
      dmin = fmin(dmin, __shfl_xor_sync(0xffffffffu, dmin, o));
    }
    use32 = dmax > 1e-290 && dmin >= dmax * 9.094947017729282e-13;  // 2^-40
    // exact power-of-two scaling: the largest diagonal entry lands in [1, 2)
    gscale = __longlong_as_double((2046LL - ((__double_as_longlong(dmax) >> 52) & 0x7ffLL)) << 52);
  }
  if (use32) {
    typedef typename LowPrec<T>::type GT;
    GT* G32 = reinterpret_cast<GT*>(G + JP * JGP);     // two single-precision buffers inside G's second buffer
    for (int idx = tid; idx < JP * JP; idx += JT) {
      const int i = idx / JP, j = idx - i * JP;
      const T v = N_::scale(G[i * JGG + j], gscale);
      if constexpr (CPLX) G32[i * JGG + j] = make_float2((float)v.x, (float)v.y); else G32[i * JGG + j] = (float)v;
    }
    __syncthreads();
    rotation_phase<T, GT>(G32, JP * JGG, W, reinterpret_cast<float (*)[JP / 2]>(s_rc),
                          reinterpret_cast<GT (*)[JP / 2]>(s_rsp), cx, cluster, state, warp, lane);
  } else {
    rotation_phase<T, T>(G, JP * JGP, W, s_rc, s_rsp, cx, cluster, state, warp, lane);
  }
  if (S > 1 && !cx.wsplit) cluster.sync();   // cluster sizes that do not split W: still nobody may leave while its partial is read
  T* const Wfree = W + ((nsteps & 1) ^ 1) * (JP * JGP);   // the buffer the last step read: free from here on
  W += (nsteps & 1) * (JP * JGP);  // the buffer the last step wrote
#else
  T* const Wfree = W + JP * JGP;
#endif
  // re + im of every entry of W, formed ONCE: the third product of the 3-multiplication form needs it as an A
  // fragment in every k-step of every 8-column block of the apply (non-tensor FP64 adds share the pipe with the
  // DMMAs: they were the largest single source of pipe-throttle stalls of the round)
  double* const Wsum = reinterpret_cast<double*>(Wfree);   // [JP][JSP]
  if constexpr (CPLX) {
    for (int idx = tid; idx < JP * JP; idx += JT) {
      const int p_ = idx / JP, q_ = idx - p_ * JP;
      const T w = W[p_ * JWP + q_];
      Wsum[p_ * JSP + q_] = w.x + w.y;
    }
    __syncthreads();
  }
  JSTAMP(75, threadIdx.x == 0);
  if (crank == 0) {
    state = __reduce_or_sync(0xffffffffu, state);
    if (lane == 0 && state) atomicOr(&flags->state, state);
  }

  // ---- rows_new[q] = sum_p W[p][q] rows[p] on every chunk of this CTA, again on DMMA: per warp a
  //      32 (q) x 16 (c) x 32 (p) product; A[m = q][k = p] = W[p][q], B[k = p][n = c] = P[p][c].
  //      A warp reads and rewrites only its own 16 columns of P, so the result goes back into P in
  //      place and is stored to global memory as full row segments.
  auto apply_chunk = [&](int gch) {
    T* base; int64_t ld, c0; int len;
    chunk_geom(gch, base, ld, c0, len);
    const bool bulk = bulk_ok(base, ld, len);
    // The chunk is processed in two halves of CH / 2 columns; inside a half a warp takes blocks of 8 columns.
    // With bulk stores the first half is on its way to global memory while the second is being computed.
    for (int h = 0; h < 2; ++h) {
      const int hbeg = h * (CH / 2), hend = hbeg + CH / 2;
      for (int cw = hbeg + warp * 8; cw < hend; cw += (JT / 32) * 8) {
        if (cw >= len) break;
        // complex: 3-multiplication form, P1 = wr pr, P2 = wi pi, P3 = (wr + wi)(pr + pi):
        //   re = P1 - P2,  im = P3 - P1 - P2   (12 DMMAs per k-step instead of 16)
        double acc[4][CPLX ? 6 : 2];
#pragma unroll
        for (int mt = 0; mt < 4; ++mt)
#pragma unroll
          for (int r = 0; r < (CPLX ? 6 : 2); ++r) acc[mt][r] = 0.0;
#ifdef TNB_EXP_SKIP_APPLY
        for (int k0 = 0; k0 < 4; k0 += 4) {
#else
#pragma unroll 2
        for (int k0 = 0; k0 < JP; k0 += 4) {
#endif
          T av[4];
#pragma unroll
          for (int mt = 0; mt < 4; ++mt) av[mt] = W[(k0 + tq) * JWP + mt * 8 + gq];
          const T bv = P[(k0 + tq) * pitch + cw + gq];
          if constexpr (CPLX) {
            const double bs = bv.x + bv.y;
#pragma unroll
            for (int mt = 0; mt < 4; ++mt) {
              dmma884(acc[mt][0], acc[mt][1], av[mt].x, bv.x);
              dmma884(acc[mt][2], acc[mt][3], av[mt].y, bv.y);
              dmma884(acc[mt][4], acc[mt][5], Wsum[(k0 + tq) * JSP + mt * 8 + gq], bs);
            }
          } else {
#pragma unroll
            for (int mt = 0; mt < 4; ++mt) dmma884(acc[mt][0], acc[mt][1], av[mt], bv);
          }
        }
        __syncwarp();
#pragma unroll
        for (int mt = 0; mt < 4; ++mt) {
          T* dst = P + (mt * 8 + gq) * pitch + cw + 2 * tq;
          if constexpr (CPLX) {
            dst[0] = make_double2(acc[mt][0] - acc[mt][2], acc[mt][4] - acc[mt][0] - acc[mt][2]);
            dst[1] = make_double2(acc[mt][1] - acc[mt][3], acc[mt][5] - acc[mt][1] - acc[mt][3]);
          } else {
            dst[0] = acc[mt][0];
            dst[1] = acc[mt][1];
          }
        }
        __syncwarp();
        if (!bulk) {
          // 32 rows x 8 columns: lane -> (row parity group lane >> 3, column lane & 7)
          const int col = cw + (lane & 7);
          if (col < len) {
#pragma unroll 4
            for (int q = (lane >> 3); q < JP; q += 4) {
              const int64_t r = grow(q);
              if (r >= 0) base[r * ld + c0 + col] = P[q * pitch + col];
            }
          }
        }
      }
      if (bulk) {
        // the half goes out as TMA bulk stores (one per row) once every warp has written its columns back into P
        fence_proxy_async_smem();
        __syncthreads();
        if (warp == 0) {
          const int64_t r = grow(lane);
          if (r >= 0) bulk_s2g(base + r * ld + c0 + hbeg, P + lane * pitch + hbeg, (uint32_t)((CH / 2) * sizeof(T)));
          bulk_commit();
          if (h == 1) bulk_wait_read();  // P may be overwritten (next chunk) or the CTA may exit after this
        }
      }
    }
  };
  if (resident >= 0) apply_chunk(resident);
  JSTAMP(76, threadIdx.x == 0);
  for (int gch = crank; gch < a.nx + a.nv; gch += S) {
    if (gch == resident) continue;
    __syncthreads();
    const bool bulk = load_chunk(gch);
    if (bulk) { wait_half(0); wait_half(1); bar_phase ^= 1; }
    __syncthreads();
    apply_chunk(gch);
  }
}

// Closes a sweep.  state bit 0: some pair was rotated (cosine above tol); bit 1: some rotated pair had a
// squared cosine above 1e-14.  Cyclic Jacobi converges quadratically (the largest cosine of a sweep is about
// the square of the previous sweep's), so a sweep whose largest cosine was already below 1e-7 leaves
// every cosine at the level of tol: the confirming sweep is skipped.
__global__ void jacobi_finish_sweep_kernel(JacobiFlags* flags, int batch) {
  griddep_wait();
  griddep_launch_dependents();
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= batch) return;
  JacobiFlags* f = flags + b;
  if (f->converged) return;
  f->sweeps += 1;
  if ((f->state & 2u) == 0u) f->converged = 1;
  f->state = 0u;
}

// sigma[j] = |Xt[j, :]|, one warp per row (scaled two-pass-free: values are O(|A|), no overflow risk
// beyond that of the input's own Frobenius norm)
template <typename T>
__global__ void __launch_bounds__(256) row_norm_kernel(const T* X, int64_t ld, int64_t n, int64_t L, double* sigma) {
  const int64_t row = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (row >= n) return;
  const int lane = threadIdx.x & 31;
  double acc = 0.0;
  for (int64_t c = lane; c < L; c += 32) acc += Num<T>::abs2(X[row * ld + c]);
  acc = warp_sum(acc);
  if (lane == 0) sigma[row] = sqrt(acc);
}

// descending rank by counting (stable): perm[rank] = j, S[rank] = sigma[j] * scale[0].
// Non-finite values rank last (so perm is always a permutation) and raise flags->bad.
__global__ void __launch_bounds__(256) rank_kernel(const double* sigma, int64_t n, int32_t* perm, double* S,
                                                   const double* scale, JacobiFlags* flags) {
  const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n) return;
  const double raw = sigma[j];
  const bool okj = isfinite(raw);
  const double sj = okj ? raw : -1.0;
  int64_t r = 0;
  for (int64_t i = 0; i < n; ++i) {
    const double ri = sigma[i];
    const double si = isfinite(ri) ? ri : -1.0;
    r += (si > sj || (si == sj && i < j)) ? 1 : 0;
  }
  perm[r] = (int32_t)j;
  const double sc = scale[0];
  S[r] = raw * sc;
  if (!okj || !isfinite(sc)) flags->bad = 1;
}

// nrm[0] = |R|_F was written by norm2; x *= 1/nrm[0] (x *= 1 for a zero or non-finite norm,
// nrm[0] is then reset to 1 so that the singular values are scaled back consistently).  A non-finite norm
// (NaN / Inf in the input) raises flags->bad: the Jacobi driver reads it with its first convergence check.
template <typename T>
__global__ void unit_scale_kernel(T* x, int64_t n, double* nrm, double* scale_out, const double* qscale,
                                  JacobiFlags* flags) {
  const double v = nrm[0];
  const bool ok = isfinite(v) && v > 0.0;
  const double inv = ok ? 1.0 / v : 1.0;
  const int64_t step = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += step) x[i] = Num<T>::scale(x[i], inv);
  // singular values of A = (row norms) * |R|_F / 2^e; a non-finite norm propagates and is reported
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    const double sc = (ok ? v : (v == 0.0 ? 1.0 : v)) * qscale[1];
    scale_out[0] = sc;
    if (!isfinite(sc)) flags->bad = 1;
  }
}

// out[k, c] = f(in[perm[k], c])            (transpose == 0, out ld = ldo)
// out[c, k] = f(in[perm[k], c])            (transpose == 1)
// f: optional conjugation and division by S[perm[k]] (S = UNSORTED row norms; zero -> zero vector)
template <typename T>
__global__ void gather_rows_kernel(const T* in, int64_t ldi, const int32_t* perm, const double* S, T* out, int64_t ldo,
                                   int64_t n, int64_t L, int conj, int scale, int transpose) {
  const int64_t step = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n * L; i += step) {
    int64_t k, c;
    if (transpose) { c = i / n; k = i - c * n; } else { k = i / L; c = i - k * L; }
    T v = in[(int64_t)perm[k] * ldi + c];
    if (conj) v = Num<T>::conj(v);
    if (scale) { const double s = S[perm[k]]; v = Num<T>::scale(v, s > 0.0 ? 1.0 / s : 0.0); }
    if (transpose) out[c * ldo + k] = v; else out[k * ldo + c] = v;
  }
}

template <typename T>
__global__ void eye_rows_kernel(T* V, int64_t n) {
  const int64_t step = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n * n; i += step) {
    const int64_t r = i / n, c = i - r * n;
    V[i] = (r == c) ? Num<T>::one() : Num<T>::zero();
  }
}

static inline size_t align_up(size_t x) { return (x + 255) & ~(size_t)255; }
static inline unsigned blocks_for(int64_t n) {
  int64_t b = (n + 255) / 256;
  const int64_t cap = (int64_t)sm_count() * 16;
  if (b > cap) b = cap;
  if (b < 1) b = 1;
  return (unsigned)b;
}

struct SvdLayout {
  int64_t k, mq, nq;  // QR problem: mq x nq with nq == k
  size_t off_ah, off_q, off_r, off_x, off_v, off_vs, off_sig, off_perm, off_flags, off_nrm, off_qr, total;
};

static SvdLayout svd_layout(int dtype, int64_t m, int64_t n) {
  SvdLayout L;
  const size_t es = elem_size(dtype);
  L.k = m < n ? m : n;
  L.mq = m < n ? n : m;
  L.nq = L.k;
  size_t o = 0;
  L.off_ah = o;    o += (m < n) ? align_up((size_t)m * n * es) : 0;
  L.off_q = o;     o += align_up((size_t)L.mq * L.k * es);
  L.off_r = o;     o += align_up((size_t)L.k * L.k * es);
  L.off_x = o;     o += align_up((size_t)L.k * L.k * es);
  L.off_v = o;     o += align_up((size_t)L.k * L.k * es);
  L.off_vs = o;    o += align_up((size_t)L.k * L.k * es);
  L.off_sig = o;   o += align_up((size_t)L.k * sizeof(double));
  L.off_perm = o;  o += align_up((size_t)L.k * sizeof(int32_t));
  L.off_flags = o; o += align_up(sizeof(JacobiFlags));
  L.off_nrm = o;   o += align_up(1024 * sizeof(double));  // [0] |R|_F, [1] scale, [8..] norm2 workspace
  L.off_qr = o;    o += align_up(qr_workspace(dtype, L.mq, L.nq));
  L.total = o;
  return L;
}

// Per-device, thread-safe cache of everything the launches need to know about the device: the opt-in for
// large dynamic shared memory (a per-device function attribute) and how many clusters of S CTAs (JT threads,
// `smem` bytes each) it can hold at once.
struct JacobiDeviceInfo {
  std::mutex mu;
  bool configured[2] = {false, false};
  int clusters[2][9][16];   // [dtype][S][smem / 16 KB], -1 = not yet queried
  int last_sweeps[2] = {0, 0};  // sweeps the previous factorisation of this dtype needed (queueing hint)
  JacobiDeviceInfo() { for (auto& d : clusters) for (auto& r : d) for (int& v : r) v = -1; }
};
static JacobiDeviceInfo& jacobi_device_info() {
  static JacobiDeviceInfo info[64];
  int dev = 0;
  cudaGetDevice(&dev);
  return info[(dev >= 0 && dev < 64) ? dev : 0];
}

template <typename T>
static int configure_round_kernel(JacobiDeviceInfo& di) {
  constexpr int dt = sizeof(T) == 16 ? 1 : 0;
  if (di.configured[dt]) return 0;
  // everything an SM offers one CTA (227 KB) minus the kernel's static shared memory
  cudaFuncAttributes fa;
  TNB_CUDA_CHECK(cudaFuncGetAttributes(&fa, jacobi_round_kernel<T>));
  TNB_CUDA_CHECK(cudaFuncSetAttribute(jacobi_round_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      227 * 1024 - (int)fa.sharedSizeBytes));
  TNB_CUDA_CHECK(cudaFuncSetAttribute(jacobi_round_kernel<T>, cudaFuncAttributeNonPortableClusterSizeAllowed, 0));
  di.configured[dt] = true;
  return 0;
}

template <typename T>
static int max_active_clusters(JacobiDeviceInfo& di, int S, size_t smem) {
  constexpr int dt = sizeof(T) == 16 ? 1 : 0;
  const int slot = (int)(smem >> 14) < 16 ? (int)(smem >> 14) : 15;
  std::lock_guard<std::mutex> lock(di.mu);
  if (di.clusters[dt][S][slot] >= 0) return di.clusters[dt][S][slot];
  if (configure_round_kernel<T>(di) != 0) { cudaGetLastError(); return 0; }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)(S * 64), 1, 1);
  cfg.blockDim = dim3(JT, 1, 1);
  cfg.dynamicSmemBytes = smem;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = (unsigned)S;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  int n = 0;
  if (cudaOccupancyMaxActiveClusters(&n, jacobi_round_kernel<T>, &cfg) != cudaSuccess) { cudaGetLastError(); n = 0; }
  di.clusters[dt][S][slot] = n;
  return n;
}

// Every launch of the Jacobi chain carries the programmatic-stream-serialization attribute: the kernels wait
// for their predecessor themselves (griddep_wait), after their prologue.
template <typename T>
static int launch_round(JacobiArgs& a, int npairs, int batch, size_t smem, cudaStream_t st) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)(npairs * a.S), (unsigned)batch, 1);
  cfg.blockDim = dim3(JT, 1, 1);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = (unsigned)a.S;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
#ifdef TNB_EXP_NO_PDL
  cfg.numAttrs = 1;  // kernel experiments: plain stream order between the rounds
#else
  cfg.numAttrs = 2;
#endif
  TNB_CUDA_CHECK(cudaLaunchKernelEx(&cfg, jacobi_round_kernel<T>, a));
  count_launch();
  return 0;
}

static int launch_finish_sweep(JacobiFlags* flags, int batch, cudaStream_t st) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)((batch + 63) / 64), 1, 1);
  cfg.blockDim = dim3(64, 1, 1);
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
#ifdef TNB_EXP_NO_PDL
  cfg.numAttrs = 0;
#else
  cfg.numAttrs = 1;
#endif
  TNB_CUDA_CHECK(cudaLaunchKernelEx(&cfg, jacobi_finish_sweep_kernel, flags, batch));
  count_launch();
  return 0;
}

// Geometry of the round launches for n columns of length L (plus n more of V when with_v).
struct JacobiPlan { int nblk, npairs, rounds, S, CH, nx, nv; size_t smem; };

template <typename T>
static JacobiPlan jacobi_plan(int64_t n, int64_t L, bool with_v, int batch) {
  JacobiDeviceInfo& di = jacobi_device_info();
  const bool cplx = sizeof(T) == 16;
  JacobiPlan p;
  const int64_t nblk_real = (n + JB - 1) / JB;
  p.nblk = (int)(nblk_real + (nblk_real & 1));
  if (p.nblk < 2) p.nblk = 2;
  p.npairs = p.nblk / 2;
  p.rounds = p.nblk - 1;
  const int ch_max = cplx ? 256 : 512;
  const int64_t longest = (L > n || !with_v) ? L : n;
  auto chunks = [&](int ch) { return (int)((L + ch - 1) / ch) + (with_v ? (int)((n + ch - 1) / ch) : 0); };
  auto smem_for = [&](int ch) { return ((size_t)JP * (ch + JPAD) + 4 * (size_t)JP * JGP + GRAM_TILES * 64) * sizeof(T); };
  // CTAs per cluster (<= 8): every pair's cluster (of every matrix of the batch) should be resident at once
  // (a second wave would double the round), so candidate sizes are checked against
  // cudaOccupancyMaxActiveClusters -- GPCs differ in SM count, so this is not simply SMs / size.  The chunk is
  // shortened until there is at least one chunk per CTA; the size with the fewest chunks per CTA wins, ties go
  // to the smaller cluster (cheaper DSMEM reduction).
  int S = 1, CH = 32, best = 1 << 30;
  while (CH < ch_max && CH < longest) CH *= 2;
  const int ch_top = CH;
  for (int c = 1; c <= 8; ++c) {
    int ch = ch_top;
    while (ch > 32 && chunks(ch) < c) ch /= 2;
    if (chunks(ch) < c) break;
    if (c > 1 && max_active_clusters<T>(di, c, smem_for(ch)) < p.npairs * batch) continue;
    const int per = (chunks(ch) + c - 1) / c;
    // elements of a row each CTA walks through, plus the price of a larger cluster (DSMEM reduction and barriers
    // over c CTAs, the rotation phase replicated c times) in the same unit: measured ~1 us ~ 32 elements per CTA
    const int cost = per * ch + 32 * c;
    if (cost < best) { best = cost; S = c; CH = ch; }
  }
  // A batch that cannot be resident at one CTA per SM: shorter chunks bring the shared memory of a CTA under
  // half an SM's, so that two CTAs (of different matrices) share an SM and one's rotation phase overlaps the
  // other's tensor-core phases.  Costs a second pass over the chunks that are not resident (L2 traffic only).
#ifndef TNB_EXP_NO_BATCH_HALVING
  if (batch > 1 && S == 1 && (int64_t)p.npairs * batch > (int64_t)sm_count()) {
    while (CH > 64 && smem_for(CH) > 110 * 1024) CH /= 2;
  }
#endif
  p.S = S; p.CH = CH;
  p.nx = (int)((L + CH - 1) / CH);
  p.nv = with_v ? (int)((n + CH - 1) / CH) : 0;
  p.smem = smem_for(CH);
#ifndef TNB_EXP_NO_SMEM_PAD
  // One matrix whose clusters all fit at one CTA per SM: ask for more than half an SM's shared memory so that the
  // block scheduler cannot put two CTAs of the round on one SM (they would walk through the same phases together
  // and share the FP64 pipe: measured 23.7 -> 19.1 us per round for k = 1024 float64, whose CTAs need only 103 KB).
  const size_t half_sm = (size_t)116 * 1024;
  if (batch == 1 && p.smem < half_sm && (int64_t)p.npairs * p.S <= (int64_t)sm_count() &&
      (p.S == 1 || max_active_clusters<T>(di, p.S, half_sm) >= p.npairs))
    p.smem = half_sm;
#endif
  return p;
}

// Orthogonalise the rows-as-columns of `batch` matrices Xt (n x L each, `xstride` elements apart) in place,
// accumulating V in Vt (n x n, must be I; may be null).  flags: one JacobiFlags per matrix, zeroed by the caller
// (a `bad` mark set before the sweeps -- non-finite input -- ends the factorisation with TNB_E_NOCONV).
//
// Host involvement: the sweeps are queued ahead of the device (a converged matrix turns its remaining
// launches into no-ops through its device flag) and the stream is synchronised only to learn whether more
// are needed -- the first batch is as long as the previous factorisation of this type needed, so that in a
// sweep over similar sites there is normally ONE synchronisation per factorisation here.
template <typename T>
static int jacobi(T* Xt, int64_t ldx, int64_t L, int64_t xstride, T* Vt, int64_t vstride, int64_t n, int batch,
                  JacobiFlags* flags, int* sweeps_out, cudaStream_t st) {
  constexpr int dt = sizeof(T) == 16 ? 1 : 0;
  const bool cplx = sizeof(T) == 16;
  JacobiDeviceInfo& di = jacobi_device_info();
  {
    std::lock_guard<std::mutex> lock(di.mu);
    int rc = configure_round_kernel<T>(di);
    if (rc) return rc;
  }
  const JacobiPlan p = jacobi_plan<T>(n, L, Vt != nullptr, batch);
  JacobiArgs a;
  a.X = Xt; a.V = Vt; a.n = n; a.L = L; a.ldx = ldx; a.ldv = n; a.nblk = p.nblk; a.flags = flags;
  a.xstride = xstride; a.vstride = vstride;
  a.tol = sqrt((double)(L > 1 ? L : 1)) * 2.220446049250313e-16;
  a.CH = p.CH; a.nx = p.nx; a.nv = p.nv; a.S = p.S;
#ifdef TNB_EXP_NO_FP32ROT
  a.fp32_rot = 0;  // kernel experiments: double-precision rotation phase everywhere
#else
  a.fp32_rot = 1;
#endif

  // conventional flop count of one sweep (full JP x JP Gram over L, W applied over L [+ n with V]; 8 real
  // flops per complex multiply-add): what the profile credits per EXECUTED sweep
  const double sweep_flops = (cplx ? 8.0 : 2.0) * (double)JP * JP * (2.0 * (double)L + (Vt ? (double)n : 0.0)) *
                             (double)p.npairs * (double)p.rounds;
  int hint;
  {
    std::lock_guard<std::mutex> lock(di.mu);
    hint = di.last_sweeps[dt];
  }
  int queued = 0, done_sweeps = 0;
  bool all_converged = false;
  std::vector<JacobiFlags> h((size_t)batch);
  while (queued < MAX_SWEEPS && !all_converged) {
    int nq = (queued == 0) ? (hint > 0 ? hint : 6) : 1;
    if (nq > MAX_SWEEPS - queued) nq = MAX_SWEEPS - queued;
    {
      ProfScope prof(KC_JACOBI, st, 0.0);
      for (int s = 0; s < nq; ++s) {
        for (int r = 0; r < p.rounds; ++r) {
          a.diag = (r == 0) ? 2 : 0;  // round 0 also rotates the pairs inside its blocks, before their cross pairs ...
          a.round = r;                // ... then every block against every other block
          int rc = launch_round<T>(a, p.npairs, batch, p.smem, st);
          if (rc) return rc;
        }
        int rc = launch_finish_sweep(flags, batch, st);
        if (rc) return rc;
      }
      queued += nq;
      TNB_CUDA_CHECK(cudaMemcpyAsync(h.data(), flags, sizeof(JacobiFlags) * (size_t)batch, cudaMemcpyDeviceToHost, st));
      TNB_CUDA_CHECK(cudaStreamSynchronize(st));
      all_converged = true;
      double executed = 0.0;   // matrix-sweeps that actually ran in this batch of launches
      int most = 0;
      for (int b = 0; b < batch; ++b) {
        if (h[(size_t)b].bad) return TNB_E_NOCONV;   // NaN / Inf in the input: numpy.linalg.svd raises LinAlgError too
        if (!h[(size_t)b].converged) all_converged = false;
        if (h[(size_t)b].sweeps > most) most = h[(size_t)b].sweeps;
        executed += (double)h[(size_t)b].sweeps;
      }
      prof.work = sweep_flops * (executed - (double)done_sweeps);
      done_sweeps = (int)executed;
      if (all_converged || queued >= MAX_SWEEPS) {
        if (sweeps_out) *sweeps_out = most;
        std::lock_guard<std::mutex> lock(di.mu);
        di.last_sweeps[dt] = most;
      }
    }
  }
  return all_converged ? 0 : TNB_E_NOCONV;
}

// mode 0: U, S, Vh (V accumulated through the Jacobi rotations).
// mode 1: U, S and P = U^H A = diag(S) Vh -- what the MPS sweeps absorb into the next site
//         (onedim_core.py:347-349 contracts V and then S into it).  V is never formed: Jacobi then
//         runs on the triangular factor only (one third fewer flops, half the traffic) and P is one
//         DMMA GEMM on the original A, accurate to eps |A| whatever the size of the singular value.
template <typename T>
static int svd_impl(int dtype, int64_t m, int64_t n, const void* A, int64_t lda, void* U, double* S, void* Vh,
                    void* ws, int* sweeps_out, int mode, cudaStream_t st) {
  const SvdLayout L = svd_layout(dtype, m, n);
  char* base = (char*)ws;
  const int64_t k = L.k;
  T* Q = (T*)(base + L.off_q);
  T* R = (T*)(base + L.off_r);
  T* Xt = (T*)(base + L.off_x);
  T* Vt = (T*)(base + L.off_v);
  T* Vs = (T*)(base + L.off_vs);
  double* sig = (double*)(base + L.off_sig);
  int32_t* perm = (int32_t*)(base + L.off_perm);
  JacobiFlags* flags = (JacobiFlags*)(base + L.off_flags);
  double* nrm = (double*)(base + L.off_nrm);
  void* qr_ws = base + L.off_qr;
  const bool wide = m < n;
  const bool proj = (mode == 1);
  int rc;
  double* qscale = nullptr;  // device: [0] power-of-two scale applied before the QR, [1] its inverse
  if (!wide) {
    rc = qr(dtype, m, n, A, lda, Q, R, qr_ws, 2, &qscale, st);
    if (rc) return rc;
  } else {
    T* AH = (T*)(base + L.off_ah);  // A^H, n x m contiguous
    const int64_t sh[2] = {n, m}, is[2] = {1, lda};
    rc = permute_view(dtype, A, 2, sh, is, AH, 1.0, 0.0, 1, st);
    if (rc) return rc;
    // projection mode never uses Q for a wide matrix (U comes from R^H, P = U^H A from A itself)
    rc = qr(dtype, n, m, AH, m, proj ? (void*)nullptr : (void*)Q, R, qr_ws, 2, &qscale, st);
    if (rc) return rc;
  }
  // Jacobi orthogonalises the columns of X, stored as rows of Xt: X = R^H (Xt = conj(R)) -- the rows of a triangular
  // factor are graded by the QR, its columns are not, and Jacobi converges on the former (Drmac-Veselic).
  // A tall (or square) matrix in projection mode wants the LEFT vectors of R without accumulating V.  Jacobi on the
  // columns of R itself delivers them (Y = R V, Ur = Y / sigma) but converges badly: 12 sweeps where R^H needs 10 on a
  // random matrix, and 24-30 where it needs 8-11 when A is rank-deficient -- the null space of R's columns only
  // emerges through cancellation, rotation by rotation (cfg 5's 4096 x 4096 boundary matrices: 23 sweeps; cfg 2's
  // numerically rank-one sites).  So R is factored once more, R^H = Q2 R2: R = R2^H Q2^H has the left vectors of
  // X = R2^H, whose columns are the graded rows of R2 (NumPy emulation, scratch/jacobi_emul.py: 30 -> 8 sweeps on a
  // rank-90 300 x 260 matrix, U an isometry over all columns to 4e-15).  Q2 is never formed.
#ifdef TNB_EXP_NO_SECOND_QR
  const bool x_is_r = proj && !wide;   // kernel experiments: Jacobi on the columns of R itself (Xt = R^T)
#else
  const bool x_is_r = false;
  if (proj && !wide) {
    const int64_t sh[2] = {k, k}, is_h[2] = {1, k};
    rc = permute_view(dtype, R, 2, sh, is_h, Vt, 1.0, 0.0, 1, st);   // Vt (free in projection mode) <- R^H
    if (rc) return rc;
    // the scale factors of the first QR live in its workspace, which the second one reuses
    TNB_CUDA_CHECK(cudaMemcpyAsync(nrm + 2, qscale, 2 * sizeof(double), cudaMemcpyDeviceToDevice, st));
    qscale = nrm + 2;
    rc = qr(dtype, k, k, Vt, k, nullptr, R, qr_ws, 0, nullptr, st);   // R <- R2
    if (rc) return rc;
  }
#endif
  {
    const int64_t sh[2] = {k, k};
    const int64_t is_conj[2] = {k, 1}, is_tr[2] = {1, k};
    rc = permute_view(dtype, R, 2, sh, x_is_r ? is_tr : is_conj, Xt, 1.0, 0.0, x_is_r ? 0 : 1, st);
    if (rc) return rc;
  }
  {  // Jacobi works on Gram matrices (squared magnitudes): bring X to unit Frobenius norm first
    tnb_tensor_t xd;
    xd.ptr = Xt; xd.dtype = dtype; xd.rank = 1; xd.shape[0] = k * k; xd.stride[0] = 1;
    rc = tnb_norm2(&xd, nrm, nrm + 8, 1016 * sizeof(double), st);
    if (rc) return rc;
    TNB_CUDA_CHECK(cudaMemsetAsync(flags, 0, sizeof(JacobiFlags), st));
    unit_scale_kernel<T><<<blocks_for(k * k), 256, 0, st>>>(Xt, k * k, nrm, nrm + 1, qscale, flags);
    TNB_LAUNCH_CHECK();
  }
  if (!proj) {
    eye_rows_kernel<T><<<blocks_for(k * k), 256, 0, st>>>(Vt, k);
    TNB_LAUNCH_CHECK();
  }
  rc = jacobi<T>(Xt, k, k, 0, proj ? (T*)nullptr : Vt, 0, k, 1, flags, sweeps_out, st);
  if (rc) return rc;
  row_norm_kernel<T><<<(unsigned)((k + 7) / 8), 256, 0, st>>>(Xt, k, k, k, sig);
  TNB_LAUNCH_CHECK();
  rank_kernel<<<(unsigned)((k + 255) / 256), 256, 0, st>>>(sig, k, perm, S, nrm + 1, flags);
  TNB_LAUNCH_CHECK();
  if (!proj) {
    // Vs[j, :] = Vt[perm[j], :]  (column j of the sorted V, as a row)
    gather_rows_kernel<T><<<blocks_for(k * k), 256, 0, st>>>(Vt, k, perm, S, Vs, k, k, k, 0, 0, 0);
    TNB_LAUNCH_CHECK();
    if (!wide) {
      // A = Q R, R = V Sigma Ux^H:  U = Q V,  Vh = Ux^H = conj(Y^T) / sigma
      if (Vh) {
        gather_rows_kernel<T><<<blocks_for(k * k), 256, 0, st>>>(Xt, k, perm, sig, (T*)Vh, n, k, k, 1, 1, 0);
        TNB_LAUNCH_CHECK();
      }
      if (U) {
        rc = gemm(dtype, TNB_OP_N, TNB_OP_T, m, k, k, 1, 0, Q, k, 0, Vs, k, 0, 0, 0, U, k, 0, 1, st);
        if (rc) return rc;
      }
    } else {
      // A = R^H Q^H = Ux Sigma (Q V)^H:  U = Y / sigma,  Vh = conj(Vs) Q^H
      if (U) {
        gather_rows_kernel<T><<<blocks_for(k * k), 256, 0, st>>>(Xt, k, perm, sig, (T*)U, k, k, k, 0, 1, 1);
        TNB_LAUNCH_CHECK();
      }
      if (Vh) {
        rc = gemm(dtype, TNB_OP_J, TNB_OP_C, k, n, k, 1, 0, Vs, k, 0, Q, k, 0, 0, 0, Vh, n, 0, 1, st);
        if (rc) return rc;
      }
    }
  } else {
    if (!U) return TNB_E_ARG;
    if (wide) {
      // A = R^H Q^H, R^H V = Y:  U = Y / sigma (left vectors of R^H are those of A)
      gather_rows_kernel<T><<<blocks_for(k * k), 256, 0, st>>>(Xt, k, perm, sig, (T*)U, k, k, k, 0, 1, 1);
      TNB_LAUNCH_CHECK();
    } else {
      // A = Q R, R = R2^H Q2^H, R2^H V = Y = Ur Sigma:  U = Q Ur, with Ur^T = sorted rows of Yt / sigma (kept in Vs)
      gather_rows_kernel<T><<<blocks_for(k * k), 256, 0, st>>>(Xt, k, perm, sig, Vs, k, k, k, 0, 1, 0);
      TNB_LAUNCH_CHECK();
      rc = gemm(dtype, TNB_OP_N, TNB_OP_T, m, k, k, 1, 0, Q, k, 0, Vs, k, 0, 0, 0, U, k, 0, 1, st);
      if (rc) return rc;
    }
    if (Vh) {  // P = U^H A  (k x n)
      rc = gemm(dtype, TNB_OP_C, TNB_OP_N, k, n, m, 1, 0, U, k, 0, A, lda, 0, 0, 0, Vh, n, 0, 1, st);
      if (rc) return rc;
    }
  }
  // No synchronisation here: non-finite input was caught before the sweeps (unit_scale_kernel -> flags->bad,
  // read with the first convergence check); finite input stays finite under unitary rotations.
  return 0;
}

// ================================================================================================
// Batched projection SVD: `batch` matrices of ONE shape in one set of launches (the batched path of
// BASELINE.json config 4: the same site of every network of a shard).  No QR pre-reduction -- at these
// sizes (a few hundred rows, chi ~ 128 columns) the whole matrix of a block pair fits the shared memory
// of one CTA, and the grid row (blockIdx.y) of the round kernel is the matrix index, so one launch per
// round serves the whole batch.  One-sided Jacobi runs directly on the columns of A (tall) or of A^H
// (wide); every matrix carries its own convergence flag.
//   tall (m >= n):  X = A      ->  Y = X V:  U = Y / sigma (sorted),  P = U^H A  (one batched GEMM)
//   wide (m <  n):  X = A^H    ->  Y = X V:  U = V (accumulated),     P = Y^H    (sorted rows)
// ================================================================================================

// One CTA per matrix: x /= |x|_F (computed as amax * |x / amax|_F: no overflow), nrm[b] = |x|_F.
// A zero matrix is left alone (nrm = 1); a non-finite one raises flags[b].bad.
template <typename T>
__global__ void __launch_bounds__(512) batched_unit_scale_kernel(T* x, int64_t per, double* nrm, JacobiFlags* flags) {
  T* xb = x + (int64_t)blockIdx.x * per;
  __shared__ double red[16];
  __shared__ double bc;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  double mx = 0.0;
  bool nan = false;
  for (int64_t i = tid; i < per; i += 512) {
    const double a = abs_t<T>(xb[i]);
    if (!(a <= 1.79e308)) nan = true;   // NaN or Inf
    mx = fmax(mx, a);
  }
  if (nan) mx = __longlong_as_double(0x7ff8000000000000LL);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const double other = __shfl_xor_sync(0xffffffffu, mx, o);
    mx = (mx != mx || other != other) ? __longlong_as_double(0x7ff8000000000000LL) : fmax(mx, other);
  }
  if (lane == 0) red[warp] = mx;
  __syncthreads();
  if (tid == 0) {
    double m_ = 0.0;
    for (int w = 0; w < 16; ++w) m_ = (m_ != m_ || red[w] != red[w]) ? red[w] + m_ : fmax(m_, red[w]);
    bc = m_;
  }
  __syncthreads();
  const double amax = bc;
  if (!(amax > 0.0) || !isfinite(amax)) {
    if (tid == 0) {
      nrm[blockIdx.x] = (amax == 0.0) ? 1.0 : amax;
      if (amax != 0.0) flags[blockIdx.x].bad = 1;
    }
    return;
  }
  const double ia = 1.0 / amax;
  double acc = 0.0;
  for (int64_t i = tid; i < per; i += 512) acc += Num<T>::abs2(Num<T>::scale(xb[i], ia));
  acc = warp_sum(acc);
  __syncthreads();
  if (lane == 0) red[warp] = acc;
  __syncthreads();
  if (tid == 0) {
    double t = 0.0;
    for (int w = 0; w < 16; ++w) t += red[w];
    bc = sqrt(t);
  }
  __syncthreads();
  const double rel = bc;                 // |x / amax|_F >= 1
  const double inv = ia / rel;
  for (int64_t i = tid; i < per; i += 512) xb[i] = Num<T>::scale(xb[i], inv);
  if (tid == 0) {
    const double v = amax * rel;
    nrm[blockIdx.x] = v;
    if (!isfinite(v)) flags[blockIdx.x].bad = 1;
  }
}

template <typename T>
__global__ void eye_rows_batched_kernel(T* V, int64_t n, int64_t total) {
  const int64_t step = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += step) {
    const int64_t r = (i / n) % n, c = i % n;
    V[i] = (r == c) ? Num<T>::one() : Num<T>::zero();
  }
}

// rank_kernel for matrix blockIdx.y: sigma / perm advance by n, S by s_stride, scale and flags by one
__global__ void __launch_bounds__(256) rank_batched_kernel(const double* sigma, int64_t n, int32_t* perm, double* S,
                                                           int64_t s_stride, const double* scale, JacobiFlags* flags) {
  const int64_t b = blockIdx.y;
  sigma += b * n; perm += b * n; S += b * s_stride;
  const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n) return;
  const double raw = sigma[j];
  const bool okj = isfinite(raw);
  const double sj = okj ? raw : -1.0;
  int64_t r = 0;
  for (int64_t i = 0; i < n; ++i) {
    const double ri = sigma[i];
    const double si = isfinite(ri) ? ri : -1.0;
    r += (si > sj || (si == sj && i < j)) ? 1 : 0;
  }
  perm[r] = (int32_t)j;
  const double sc = scale[b];
  S[r] = raw * sc;
  if (!okj || !isfinite(sc)) flags[b].bad = 1;
}

// gather_rows_kernel for matrix blockIdx.y (in / out advance by their strides, perm and S by n);
// `mul` != null multiplies by mul[b] (undoes the unit-norm scaling)
template <typename T>
__global__ void gather_rows_batched_kernel(const T* in, int64_t ldi, int64_t in_stride, const int32_t* perm, const double* S,
                                           T* out, int64_t ldo, int64_t out_stride, int64_t n, int64_t L, int conj, int scale,
                                           int transpose, const double* mul) {
  const int64_t b = blockIdx.y;
  in += b * in_stride; out += b * out_stride; perm += b * n; S += b * n;
  const double f = mul ? mul[b] : 1.0;
  const int64_t step = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n * L; i += step) {
    int64_t k, c;
    if (transpose) { c = i / n; k = i - c * n; } else { k = i / L; c = i - k * L; }
    T v = in[(int64_t)perm[k] * ldi + c];
    if (conj) v = Num<T>::conj(v);
    double g = f;
    if (scale) { const double s_ = S[perm[k]]; g = s_ > 0.0 ? f / s_ : 0.0; }
    v = Num<T>::scale(v, g);
    if (transpose) out[c * ldo + k] = v; else out[k * ldo + c] = v;
  }
}

struct SvdBatchLayout { int64_t k, Lx; size_t off_x, off_v, off_sig, off_perm, off_flags, off_nrm, total; };
static SvdBatchLayout svd_batch_layout(int dtype, int64_t m, int64_t n, int64_t batch) {
  SvdBatchLayout L;
  const size_t es = elem_size(dtype);
  const bool wide = m < n;
  L.k = wide ? m : n;
  L.Lx = wide ? n : m;
  size_t o = 0;
  L.off_x = o;     o += align_up((size_t)batch * m * n * es);
  L.off_v = o;     o += wide ? align_up((size_t)batch * L.k * L.k * es) : 0;
  L.off_sig = o;   o += align_up((size_t)batch * L.k * sizeof(double));
  L.off_perm = o;  o += align_up((size_t)batch * L.k * sizeof(int32_t));
  L.off_flags = o; o += align_up((size_t)batch * sizeof(JacobiFlags));
  L.off_nrm = o;   o += align_up((size_t)batch * sizeof(double));
  L.total = o;
  return L;
}

template <typename T>
static int svd_project_batched_impl(int dtype, int64_t m, int64_t n, int64_t batch, const void* A, int64_t lda, int64_t sA,
                                    void* U, int64_t sU, double* S, int64_t sS, void* P, int64_t sP, void* ws,
                                    int* sweeps_out, cudaStream_t st) {
  const SvdBatchLayout L = svd_batch_layout(dtype, m, n, batch);
  char* base = (char*)ws;
  const bool wide = m < n;
  const int64_t k = L.k, Lx = L.Lx;
  T* Xt = (T*)(base + L.off_x);
  T* Vt = wide ? (T*)(base + L.off_v) : nullptr;
  double* sig = (double*)(base + L.off_sig);
  int32_t* perm = (int32_t*)(base + L.off_perm);
  JacobiFlags* flags = (JacobiFlags*)(base + L.off_flags);
  double* nrm = (double*)(base + L.off_nrm);
  int rc;
  {  // Xt[b] = A[b]^T (tall: the columns of A become contiguous rows) or conj(A[b]) (wide: X = A^H)
    const int64_t sh[3] = {batch, k, Lx};
    const int64_t is_tall[3] = {sA, 1, lda}, is_wide[3] = {sA, lda, 1};
    rc = permute_view(dtype, A, 3, sh, wide ? is_wide : is_tall, Xt, 1.0, 0.0, wide ? 1 : 0, st);
    if (rc) return rc;
  }
  TNB_CUDA_CHECK(cudaMemsetAsync(flags, 0, sizeof(JacobiFlags) * (size_t)batch, st));
  batched_unit_scale_kernel<T><<<(unsigned)batch, 512, 0, st>>>(Xt, k * Lx, nrm, flags);
  TNB_LAUNCH_CHECK();
  if (wide) {
    eye_rows_batched_kernel<T><<<blocks_for(batch * k * k), 256, 0, st>>>(Vt, k, batch * k * k);
    TNB_LAUNCH_CHECK();
  }
  rc = jacobi<T>(Xt, Lx, Lx, k * Lx, Vt, k * k, k, (int)batch, flags, sweeps_out, st);
  if (rc) return rc;
  row_norm_kernel<T><<<(unsigned)((batch * k + 7) / 8), 256, 0, st>>>(Xt, Lx, batch * k, Lx, sig);
  TNB_LAUNCH_CHECK();
  rank_batched_kernel<<<dim3((unsigned)((k + 255) / 256), (unsigned)batch), 256, 0, st>>>(sig, k, perm, S, sS, nrm, flags);
  TNB_LAUNCH_CHECK();
  unsigned gx = blocks_for(k * Lx);
  if (gx > 64) gx = 64;
  if (!wide) {
    gather_rows_batched_kernel<T><<<dim3(gx, (unsigned)batch), 256, 0, st>>>(Xt, Lx, k * Lx, perm, sig, (T*)U, k, sU, k, Lx, 0,
                                                                          1, 1, nullptr);
    TNB_LAUNCH_CHECK();
    if (P) {
      rc = gemm(dtype, TNB_OP_C, TNB_OP_N, k, n, m, 1, 0, U, k, sU, A, lda, sA, 0, 0, P, n, sP, batch, st);
      if (rc) return rc;
    }
  } else {
    if (P) {
      gather_rows_batched_kernel<T><<<dim3(gx, (unsigned)batch), 256, 0, st>>>(Xt, Lx, k * Lx, perm, sig, (T*)P, n, sP, k, Lx,
                                                                            1, 0, 0, nrm);
      TNB_LAUNCH_CHECK();
    }
    gather_rows_batched_kernel<T><<<dim3(gx, (unsigned)batch), 256, 0, st>>>(Vt, k, k * k, perm, sig, (T*)U, k, sU, k, k, 0, 0,
                                                                          1, nullptr);
    TNB_LAUNCH_CHECK();
  }
  return 0;
}

}  // namespace tnb

#ifdef TNB_EXP_STAMPS
extern "C" int tnb_debug_jacobi_stamps(long long* out, int n) {
  if (!out || n <= 0 || n > 128) return TNB_E_ARG;
  TNB_CUDA_CHECK(cudaMemcpyFromSymbol(out, tnb::g_jac_dbg, (size_t)n * sizeof(long long)));
  return 0;
}
#endif

extern "C" size_t tnb_svd_workspace(int dtype, int64_t m, int64_t n) {
  if (m <= 0 || n <= 0 || (dtype != TNB_F64 && dtype != TNB_C128)) return 0;
  return tnb::svd_layout(dtype, m, n).total;
}

extern "C" int tnb_svd(int dtype, int64_t m, int64_t n, const void* A, int64_t lda, void* U, double* S, void* Vh,
                       void* ws, size_t ws_bytes, int* sweeps_out, void* stream) {
  if (dtype != TNB_F64 && dtype != TNB_C128) return TNB_E_ARG;
  if (m < 0 || n < 0 || lda < n) return TNB_E_ARG;
  if (sweeps_out) *sweeps_out = 0;
  if (m == 0 || n == 0) return 0;
  if (!A || !S || !ws) return TNB_E_ARG;
  if (ws_bytes < tnb::svd_layout(dtype, m, n).total) return TNB_E_WORKSPACE;
  if (dtype == TNB_F64) return tnb::svd_impl<double>(dtype, m, n, A, lda, U, S, Vh, ws, sweeps_out, 0, (cudaStream_t)stream);
  return tnb::svd_impl<tnb::cplx>(dtype, m, n, A, lda, U, S, Vh, ws, sweeps_out, 0, (cudaStream_t)stream);
}

extern "C" int tnb_svd_project(int dtype, int64_t m, int64_t n, const void* A, int64_t lda, void* U, double* S, void* P,
                               void* ws, size_t ws_bytes, int* sweeps_out, void* stream) {
  if (dtype != TNB_F64 && dtype != TNB_C128) return TNB_E_ARG;
  if (m < 0 || n < 0 || lda < n) return TNB_E_ARG;
  if (sweeps_out) *sweeps_out = 0;
  if (m == 0 || n == 0) return 0;
  if (!A || !S || !U || !ws) return TNB_E_ARG;
  if (ws_bytes < tnb::svd_layout(dtype, m, n).total) return TNB_E_WORKSPACE;
  if (dtype == TNB_F64) return tnb::svd_impl<double>(dtype, m, n, A, lda, U, S, P, ws, sweeps_out, 1, (cudaStream_t)stream);
  return tnb::svd_impl<tnb::cplx>(dtype, m, n, A, lda, U, S, P, ws, sweeps_out, 1, (cudaStream_t)stream);
}

extern "C" size_t tnb_svd_project_batched_workspace(int dtype, int64_t m, int64_t n, int64_t batch) {
  if (m <= 0 || n <= 0 || batch <= 0 || (dtype != TNB_F64 && dtype != TNB_C128)) return 0;
  return tnb::svd_batch_layout(dtype, m, n, batch).total;
}

extern "C" int tnb_svd_project_batched(int dtype, int64_t m, int64_t n, int64_t batch, const void* A, int64_t lda,
                                       int64_t strideA, void* U, int64_t strideU, double* S, int64_t strideS, void* P,
                                       int64_t strideP, void* ws, size_t ws_bytes, int* sweeps_out, void* stream) {
  if (dtype != TNB_F64 && dtype != TNB_C128) return TNB_E_ARG;
  if (m < 0 || n < 0 || batch < 0 || lda < n) return TNB_E_ARG;
  if (sweeps_out) *sweeps_out = 0;
  if (m == 0 || n == 0 || batch == 0) return 0;
  if (batch > 65535) return TNB_E_UNSUPPORTED;
  const int64_t k = m < n ? m : n;
  if (!A || !S || !U || !ws || strideA < m * lda || strideU < m * k || strideS < k || (P && strideP < k * n)) return TNB_E_ARG;
  if (ws_bytes < tnb::svd_batch_layout(dtype, m, n, batch).total) return TNB_E_WORKSPACE;
  if (dtype == TNB_F64)
    return tnb::svd_project_batched_impl<double>(dtype, m, n, batch, A, lda, strideA, U, strideU, S, strideS, P, strideP, ws,
                                                 sweeps_out, (cudaStream_t)stream);
  return tnb::svd_project_batched_impl<tnb::cplx>(dtype, m, n, batch, A, lda, strideA, U, strideU, S, strideS, P, strideP, ws,
                                                  sweeps_out, (cudaStream_t)stream);
}
