// Reduced QR factorisation (np.linalg.qr(mode="reduced"), reference call site
// tensor.py:1044) for float64 / complex128: A = Q R, Q with orthonormal columns,
// R upper triangular (trapezoidal for m < n) with a REAL diagonal.
//
// Two paths, chosen in qr():
//   * k = min(m, n) >= 128: block Gram-Schmidt with reorthogonalisation
//     (BCGS-PIP+, qr_bcgs2 below): every O(m n^2) operation is a DMMA GEMM over
//     all SMs, the only serial kernel is a 64 x 64 Cholesky + triangular inverse
//     per column block; the second pass is lagged to groups of 256 columns and
//     needs no Cholesky at all (first-order factor of a near-identity Gram
//     matrix); device checks fall back to the path below.
//   * otherwise, and as the fallback: blocked Householder with compact-WY
//     updates (LAPACK geqrf/ungqr conventions: H_j = I - tau_j v_j v_j^H,
//     R = H_k^H ... H_1^H A, Q = H_1 ... H_k [I; 0]):
//       - panel (NB = 32 columns): ONE thread-block cluster owns the whole panel
//         in shared memory, rows split across the CTAs of the cluster
//         (CholeskyQR2 + Householder reconstruction, or column-by-column
//         Householder with one DSMEM all-reduce per column for ill-conditioned
//         panels).  The panel also emits V (explicit, unit lower trapezoid) and
//         the triangular factor T of the block reflector.
//       - trailing matrix / explicit Q: three GEMMs per 128-column outer block on
//         the FP64 tensor pipe (gemm.cu): W = V^H A2, W = T^H W, A2 -= V W.
// Nominal flops (LAPACK model): 2(2mn^2 - 2/3 n^3) real, x4 complex.
#include <cooperative_groups.h>
#include <stdlib.h>

#include "common.cuh"

namespace cg = cooperative_groups;

namespace tnb {

int gemm(int dtype, int opA, int opB, int64_t M, int64_t N, int64_t K, double ar, double ai, const void* A,
         int64_t lda, int64_t sA, const void* B, int64_t ldb, int64_t sB, double br, double bi, void* C, int64_t ldc,
         int64_t sC, int64_t batch, cudaStream_t st);
int gemm_ws(int dtype, int opA, int opB, int64_t M, int64_t N, int64_t K, double ar, double ai, const void* A,
            int64_t lda, int64_t sA, const void* B, int64_t ldb, int64_t sB, double br, double bi, void* C, int64_t ldc,
            int64_t sC, int64_t batch, void* splitk_ws, size_t splitk_bytes, cudaStream_t st);

constexpr int QR_NB = 32;
constexpr int QR_THREADS = 256;
constexpr int QR_WARPS = QR_THREADS / 32;

struct QrPanelArgs {
  void* W;        // working copy of A, m x n, ld = ldw
  void* V;        // explicit Householder vectors, m x k, ld = ldv
  void* T;        // this panel's jb x jb triangular factor (row-major, ld = ldt)
  int64_t m, ldw, ldv, ldt;
  int64_t j0;     // first column / diagonal row of the panel
  int jb;         // panel width (<= QR_NB)
  int fast;       // try the CholeskyQR2 + Householder-reconstruction path first
  int rows_per;   // panel rows owned by each CTA of the cluster
  int pitch;      // shared-memory pitch of one panel column (elements)
};

template <typename T> __device__ __forceinline__ T warp_sum_t(T v);
template <> __device__ __forceinline__ double warp_sum_t<double>(double v) { return warp_sum(v); }
template <> __device__ __forceinline__ cplx warp_sum_t<cplx>(cplx v) {
  return make_double2(warp_sum(v.x), warp_sum(v.y));
}

// kernel experiments (build with -DTNB_EXP_STAMPS -DTNB_EXP_QR_FAST_PANEL=3): phase timestamps of CTA 0 /
// thread 0 of the LAST panel launch.  Release builds compile none of this.
#ifdef TNB_EXP_STAMPS
__device__ long long g_qr_dbg[32];
#define QR_STAMP(k) do { if ((a.fast & 2) && crank == 0 && tid == 0) g_qr_dbg[k] = clock64(); } while (0)
#else
#define QR_STAMP(k) do { } while (0)
#endif

constexpr int QP = QR_NB + 4;  // pitch of the 32 x 32 matrices of the fast path (conflict-free DMMA fragments)

// G (32 x 32, pitch QP) += this CTA's P^H P over K4 rows, on DMMA; 8 warps, two 8 x 8 tiles each.
template <typename T>
__device__ __forceinline__ void panel_gram(const T* P, int pitch, int K4, T* G, int warp, int lane) {
  constexpr bool CPLX = (sizeof(T) == 16);
  const int gq = lane >> 2, tq = lane & 3;
  const int gm = warp >> 1, gn0 = (warp & 1) * 2;
  double g[2][CPLX ? 4 : 2];
#pragma unroll
  for (int j = 0; j < 2; ++j)
#pragma unroll
    for (int r = 0; r < (CPLX ? 4 : 2); ++r) g[j][r] = 0.0;
  const T* pa = P + (gm * 8 + gq) * pitch + tq;
  const T* pb0 = P + (gn0 * 8 + gq) * pitch + tq;
  const T* pb1 = pb0 + 8 * pitch;
#pragma unroll 4
  for (int k0 = 0; k0 < K4; k0 += 4) {
    const T av = pa[k0], b0 = pb0[k0], b1 = pb1[k0];
    if constexpr (CPLX) {
      const double nay = -av.y;
      dmma884(g[0][0], g[0][1], av.x, b0.x);
      dmma884(g[0][2], g[0][3], av.x, b0.y);
      dmma884(g[1][0], g[1][1], av.x, b1.x);
      dmma884(g[1][2], g[1][3], av.x, b1.y);
      dmma884(g[0][0], g[0][1], av.y, b0.y);
      dmma884(g[0][2], g[0][3], nay, b0.x);
      dmma884(g[1][0], g[1][1], av.y, b1.y);
      dmma884(g[1][2], g[1][3], nay, b1.x);
    } else {
      dmma884(g[0][0], g[0][1], av, b0);
      dmma884(g[1][0], g[1][1], av, b1);
    }
  }
#pragma unroll
  for (int j = 0; j < 2; ++j) {
    const int r = gm * 8 + gq, c = (gn0 + j) * 8 + 2 * tq;
    if constexpr (CPLX) {
      G[r * QP + c] = make_double2(g[j][0], g[j][2]);
      G[r * QP + c + 1] = make_double2(g[j][1], g[j][3]);
    } else {
      G[r * QP + c] = g[j][0];
      G[r * QP + c + 1] = g[j][1];
    }
  }
}

// In-place Cholesky G = R^H R of a 32 x 32 Hermitian matrix (pitch QP) by the whole CTA (256 threads):
// on return the upper triangle holds R (real positive diagonal), invd[k] = 1 / R[k][k].  Returns false
// (for every thread alike) when a pivot falls below `rel_floor` times its original diagonal entry --
// the panel is then too ill-conditioned for CholeskyQR and the caller takes the Householder path.
template <typename T>
__device__ __forceinline__ bool panel_cholesky(T* G, T* invd, double* diag0, double rel_floor, int tid) {
  typedef Num<T> N_;
  const int i0 = (tid >> 5) * 4, k = tid & 31;  // this thread owns rows i0..i0+3 of column k
  if (tid < QR_NB) diag0[tid] = N_::real(G[tid * QP + tid]);
  bool ok = true;
  for (int j = 0; j < QR_NB; ++j) {
    __syncthreads();
    const double piv = N_::real(G[j * QP + j]);
    if (!(piv > rel_floor * diag0[j]) || !(piv > 0.0)) ok = false;
    const double ipiv = __drcp_rn(piv);
    const T rjk = G[j * QP + k];
#pragma unroll
    for (int ii = 0; ii < 4; ++ii) {
      const int i = i0 + ii;
      if (i > j && k >= i) {
        const T f = N_::scale(N_::conj(G[j * QP + i]), ipiv);
        G[i * QP + k] = N_::sub(G[i * QP + k], N_::mul(f, rjk));
      }
    }
  }
  __syncthreads();
  // rows were kept unscaled (Schur complements): R[j][k] = G[j][k] / sqrt(G[j][j])
  double rs[4];
#pragma unroll
  for (int ii = 0; ii < 4; ++ii) rs[ii] = rsqrt(N_::real(G[(i0 + ii) * QP + i0 + ii]));
  __syncthreads();
#pragma unroll
  for (int ii = 0; ii < 4; ++ii) {
    const int i = i0 + ii;
    T v = (k >= i) ? N_::scale(G[i * QP + k], rs[ii]) : N_::zero();
    if (k == i) { v = N_::from(N_::real(v), 0.0); invd[i] = N_::from(rs[ii], 0.0); }  // exactly real diagonal
    G[i * QP + k] = v;
  }
  __syncthreads();
  return ok;
}

// Minv = M^-1 for a 32 x 32 upper triangular M (pitch QP) by recursive doubling, whole CTA (256 threads):
// the diagonal first, then for block sizes b = 1, 2, .. 16 the off-diagonal blocks
// X12 = -A^-1 B C^-1 of each pair of adjacent diagonal blocks.  `tmp` is a QR_NB x QP scratch matrix.
// Serial depth ~ 2 * 31 complex multiply-adds instead of the ~500 of a column-by-column substitution.
template <typename T>
__device__ __forceinline__ void panel_tri_inverse(const T* M, T* Minv, T* tmp, int tid) {
  typedef Num<T> N_;
  for (int idx = tid; idx < QR_NB * QR_NB; idx += QR_THREADS) {
    const int i = idx / QR_NB, j = idx - i * QR_NB;
    T v = N_::zero();
    if (i == j) {
      const T u = M[i * QP + i];
      v = N_::scale(N_::conj(u), 1.0 / N_::abs2(u));
    }
    Minv[i * QP + j] = v;
  }
  __syncthreads();
  for (int b = 1; b < QR_NB; b *= 2) {
    // entries of all off-diagonal blocks at this level: (QR_NB / 2b) blocks of b x b = QR_NB * b / 2 entries
    const int nent = QR_NB * b / 2;
    // tmp = B Cinv   (B = M[s..s+b, s+b..s+2b), Cinv = Minv[s+b.., s+b..])
    for (int e = tid; e < nent; e += QR_THREADS) {
      const int blk = e / (b * b), r = (e / b) % b, c = e % b;
      const int s0 = blk * 2 * b;
      T sum = N_::zero();
      for (int q = 0; q <= c; ++q) sum = N_::fma(M[(s0 + r) * QP + s0 + b + q], Minv[(s0 + b + q) * QP + s0 + b + c], sum);
      tmp[(s0 + r) * QP + s0 + b + c] = sum;
    }
    __syncthreads();
    // X12 = -Ainv tmp
    for (int e = tid; e < nent; e += QR_THREADS) {
      const int blk = e / (b * b), r = (e / b) % b, c = e % b;
      const int s0 = blk * 2 * b;
      T sum = N_::zero();
      for (int q = r; q < b; ++q) sum = N_::fma(Minv[(s0 + r) * QP + s0 + q], tmp[(s0 + q) * QP + s0 + b + c], sum);
      Minv[(s0 + r) * QP + s0 + b + c] = N_::sub(N_::zero(), sum);
    }
    __syncthreads();
  }
}

// P (rows r_begin..nrows of the slab) <- P * B for a 32 x 32 matrix B (pitch QP), on DMMA.  Each warp owns
// whole 8-row tiles: it reads every column of its rows before writing them back, so the product is in place.
template <typename T>
__device__ __forceinline__ void panel_right_mult(T* P, int pitch, int r_begin, int nrows, const T* B, int warp,
                                                 int lane) {
  constexpr bool CPLX = (sizeof(T) == 16);
  typedef Num<T> N_;
  const int gq = lane >> 2, tq = lane & 3;
  for (int m0 = r_begin + warp * 8; m0 < nrows; m0 += QR_WARPS * 8) {
    double acc[4][CPLX ? 4 : 2];
#pragma unroll
    for (int nt = 0; nt < 4; ++nt)
#pragma unroll
      for (int r = 0; r < (CPLX ? 4 : 2); ++r) acc[nt][r] = 0.0;
    const int row = m0 + gq;
    // lanes whose row lies past the slab load the tile's last valid row instead (their result is never stored):
    // a row index past the pitch would alias the next column's first rows, which another warp is rewriting
    // (racecheck, round 2) -- harmless for the stored rows, but a read of memory in flux all the same
    const int lrow = row < nrows ? row : nrows - 1;
#pragma unroll 2
    for (int k0 = 0; k0 < QR_NB; k0 += 4) {
      const T av = P[(k0 + tq) * pitch + lrow];
#pragma unroll
      for (int nt = 0; nt < 4; ++nt) {
        const T bv = B[(k0 + tq) * QP + nt * 8 + gq];
        if constexpr (CPLX) {
          dmma884(acc[nt][0], acc[nt][1], av.x, bv.x);
          dmma884(acc[nt][2], acc[nt][3], av.x, bv.y);
          dmma884(acc[nt][0], acc[nt][1], -av.y, bv.y);
          dmma884(acc[nt][2], acc[nt][3], av.y, bv.x);
        } else {
          dmma884(acc[nt][0], acc[nt][1], av, bv);
        }
      }
    }
    __syncwarp();
    if (row < nrows) {
#pragma unroll
      for (int nt = 0; nt < 4; ++nt) {
        const int c = nt * 8 + 2 * tq;
        if constexpr (CPLX) {
          P[c * pitch + row] = make_double2(acc[nt][0], acc[nt][2]);
          P[(c + 1) * pitch + row] = make_double2(acc[nt][1], acc[nt][3]);
        } else {
          P[c * pitch + row] = acc[nt][0];
          P[(c + 1) * pitch + row] = acc[nt][1];
        }
      }
    }
    __syncwarp();
  }
}

// Shared-memory layout (dynamic):
//   P    [jb][pitch]      panel slab, column-major (column c contiguous over the CTA's rows)
//   part [2][QR_NB]       this CTA's partial dots, double-buffered by column parity
//   rowv [2][QR_NB]       the diagonal row A[j, c] (written by the CTA that owns row j)
//   tau  [QR_NB]
//   Z    [QR_NB][QR_NB]   partial V^H V, later T
template <typename T>
__global__ void __launch_bounds__(QR_THREADS) qr_panel_kernel(QrPanelArgs a) {
  typedef Num<T> N_;
  cg::cluster_group cluster = cg::this_cluster();
  const int C = (int)cluster.num_blocks();
  const int crank = (int)cluster.block_rank();
  extern __shared__ __align__(16) unsigned char smem_raw[];
  T* P = reinterpret_cast<T*>(smem_raw);
  T* part = P + (size_t)QR_NB * a.pitch;
  T* rowv = part + 2 * QR_NB;
  T* tau = rowv + 2 * QR_NB;
  T* Z = tau + QR_NB;
  T* SA = Z + QR_NB * QR_NB;  // [QR_NB][QP]  fast path: Gram partial / top block of Q / L and U
  T* SB = SA + QR_NB * QP;    // [QR_NB][QP]  fast path: Gram total -> Cholesky factor / (U R2)^-1
  T* SC = SB + QR_NB * QP;    // [QR_NB][QP]  fast path: scratch / top block of Q -> L and U

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int jb = a.jb;
  const int64_t mr = a.m - a.j0;                 // rows in the panel
  const int64_t r_lo = (int64_t)crank * a.rows_per;  // first panel row of this CTA
  int nrows = (int)((mr - r_lo < a.rows_per) ? (mr - r_lo) : a.rows_per);
  if (nrows < 0) nrows = 0;
  T* Wg = reinterpret_cast<T*>(a.W);
  T* Vg = reinterpret_cast<T*>(a.V);

  QR_STAMP(0);
  // ---- load the slab (global rows are contiguous along c) --------------------------
  for (int i = tid; i < nrows * jb; i += QR_THREADS) {
    const int r = i / jb, c = i - r * jb;
    P[c * a.pitch + r] = Wg[(a.j0 + r_lo + r) * a.ldw + a.j0 + c];
  }
  __syncthreads();

  // =====================================================================================================
  // Fast path (full 32-column panels whose first 32 rows sit in CTA 0): CholeskyQR2 + Householder
  // reconstruction.  Q R = panel by two rounds of  G = P^H P (DMMA, one cluster reduction),  G = R^H R,
  // P <- P R^-1;  then the Householder form of that Q (Ballard, Demmel, Grigori, Jacquelin, Nguyen,
  // Solomonik, "Reconstructing Householder vectors from tall-skinny QR"):  Q - [S; 0] = L U without
  // pivoting with S = -sign(Re diag) chosen on the fly (pivots have modulus >= 1),  V = L,
  // T = -U S L1^-H,  R_householder = S R.  Six cluster barriers instead of one per column.  A panel whose
  // Cholesky pivots fall below 1e-6 of the squared column norms (or a second pass that is not close to I) is
  // too ill-conditioned for this and drops to the column-by-column Householder loop below, reloaded.
  // =====================================================================================================
  if (jb == QR_NB && a.rows_per >= QR_NB && a.fast) {
    __shared__ double diag0[QR_NB];
    __shared__ double sgn[QR_NB];
    __shared__ T invd[QR_NB];
    const int K4f = (nrows + 3) & ~3;
    for (int i = tid; i < (K4f - nrows) * QR_NB; i += QR_THREADS) {
      const int c = i / (K4f - nrows), r = nrows + (i - c * (K4f - nrows));
      P[c * a.pitch + r] = N_::zero();
    }
    __syncthreads();
    bool ok = true;
    QR_STAMP(1);
    for (int pass = 0; pass < 2 && ok; ++pass) {
      panel_gram<T>(P, a.pitch, K4f, SA, warp, lane);
      QR_STAMP(2 + 6 * pass);
      cluster.sync();
      // all-reduce over the cluster as reduce-scatter + all-gather through distributed shared memory: CTA c
      // sums entries [c * per, (c + 1) * per) of all partials (in rank order: every entry is summed exactly
      // once, so all CTAs end up with identical bits), then every CTA collects the 1024 sums.
      {
        const int per = (QR_NB * QR_NB) / C;  // C is a power of two <= 16
        T* red = SC;                          // per (<= 1024) sums of this CTA
        for (int e = tid; e < per; e += QR_THREADS) {
          const int idx = crank * per + e, i = idx / QR_NB, j = idx - i * QR_NB;
          T sum = N_::zero();
          for (int q = 0; q < C; ++q) sum = N_::add(sum, cluster.map_shared_rank(SA, q)[i * QP + j]);
          red[e] = sum;
        }
        cluster.sync();
        for (int idx = tid; idx < QR_NB * QR_NB; idx += QR_THREADS) {
          const int i = idx / QR_NB, j = idx - i * QR_NB;
          SB[i * QP + j] = cluster.map_shared_rank(red, idx / per)[idx % per];
        }
      }
      cluster.sync();  // every CTA has read every partial / sum before SA and SC are reused
      QR_STAMP(3 + 6 * pass);
      ok = panel_cholesky<T>(SB, invd, diag0, pass == 0 ? 1e-6 : 0.25, tid);  // SB <- R (upper)
      QR_STAMP(4 + 6 * pass);
      if (!ok) break;
      if (pass == 0) {
        for (int idx = tid; idx < QR_NB * QR_NB; idx += QR_THREADS) Z[idx] = SB[(idx / QR_NB) * QP + (idx % QR_NB)];
        panel_tri_inverse<T>(SB, SA, SC, tid);                 // SA <- R1^-1
        QR_STAMP(5);
        panel_right_mult<T>(P, a.pitch, 0, nrows, SA, warp, lane);  // P <- P R1^-1 = Q1
        __syncthreads();
        QR_STAMP(6);
      }
    }
    if (ok) {
      // R_total = R2 R1 (kept in Z, row-major pitch QR_NB): thread (i, k) with a dot product over q = i..k
      {
        T newz[4];
        const int i0 = (tid >> 5) * 4, k = tid & 31;
#pragma unroll
        for (int ii = 0; ii < 4; ++ii) {
          const int i = i0 + ii;
          T sum = N_::zero();
          for (int q = i; q <= k; ++q) sum = N_::fma(SB[i * QP + q], Z[q * QR_NB + k], sum);
          newz[ii] = sum;
        }
        __syncthreads();
#pragma unroll
        for (int ii = 0; ii < 4; ++ii) Z[(i0 + ii) * QR_NB + k] = newz[ii];
        __syncthreads();
      }
      // ---- Householder reconstruction on CTA 0 (Q = Q1 R2^-1 is never formed below the top block) -----
      QR_STAMP(12);
      if (crank == 0) {
        panel_tri_inverse<T>(SB, SA, SC, tid);  // SA <- R2^-1
        {  // SC <- Q_top = Q1_top R2^-1
          const int i = tid >> 3, k0 = (tid & 7) * 4;
          T out[4];
#pragma unroll
          for (int kk = 0; kk < 4; ++kk) {
            const int k = k0 + kk;
            T sum = N_::zero();
            for (int q = 0; q <= k; ++q) sum = N_::fma(P[q * a.pitch + i], SA[q * QP + k], sum);
            out[kk] = sum;
          }
#pragma unroll
          for (int kk = 0; kk < 4; ++kk) SC[i * QP + k0 + kk] = out[kk];
        }
        QR_STAMP(13);
        // LU of Q_top - S without pivoting, S_jj = -sign(Re pivot candidate) chosen on the fly
        const int i = tid >> 3, k0 = (tid & 7) * 4;  // thread owns SC[i][k0..k0+3]
        __shared__ T lu_ipiv[QR_NB];
        for (int j = 0; j < QR_NB; ++j) {
          __syncthreads();
          // step j reads row j and column j and writes only rows > j, columns > j: one barrier per step.
          // Column j keeps the unscaled a_ij; L[i][j] = a_ij / piv_j is formed after the loop.
          const T d = SC[j * QP + j];
          const double sj = (N_::real(d) >= 0.0) ? -1.0 : 1.0;
          const T piv = N_::sub(d, N_::from(sj, 0.0));
          const T ipiv = N_::scale(N_::conj(piv), __drcp_rn(N_::abs2(piv)));
          if (tid == j * 8) { lu_ipiv[j] = ipiv; sgn[j] = sj; }
          if (i > j) {
            const T lij = N_::mul(SC[i * QP + j], ipiv);
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) {
              const int k = k0 + kk;
              if (k > j) SC[i * QP + k] = N_::sub(SC[i * QP + k], N_::mul(lij, SC[j * QP + k]));
            }
          }
        }
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) {
          const int k = k0 + kk;
          if (i > k) SC[i * QP + k] = N_::mul(SC[i * QP + k], lu_ipiv[k]);            // L
          else if (i == k) SC[i * QP + k] = N_::sub(SC[i * QP + k], N_::from(sgn[k], 0.0));  // U diagonal = pivot
        }
        __syncthreads();
        QR_STAMP(14);
        // T = -U S L1^-H: row i solves  t L1^H = -(U S)[i][:]  (L1^H unit upper triangular); built in the
        // upper triangle of SA (R2^-1 is no longer needed)
        if (tid < QR_NB) {
          const int ti = tid;
          for (int k = ti; k < QR_NB; ++k) {
            T acc = N_::scale(SC[ti * QP + k], -sgn[k]);
            for (int q = ti; q < k; ++q) acc = N_::sub(acc, N_::mul(SA[ti * QP + q], N_::conj(SC[k * QP + q])));
            SA[ti * QP + k] = acc;
          }
        }
        __syncthreads();
        T* Tg = reinterpret_cast<T*>(a.T);
        for (int idx = tid; idx < QR_NB * QR_NB; idx += QR_THREADS) {
          const int r = idx / QR_NB, k = idx - r * QR_NB;
          Tg[r * a.ldt + k] = (k >= r) ? SA[r * QP + k] : N_::zero();
          // R_householder = S R_total into the working matrix
          if (k >= r) Wg[(a.j0 + r) * a.ldw + a.j0 + k] = N_::scale(Z[r * QR_NB + k], sgn[r]);
          // top block of V = L1 (unit lower triangular); Q1_top is no longer needed
          P[k * a.pitch + r] = (r > k) ? SC[r * QP + k] : (r == k ? N_::one() : N_::zero());
        }
        __syncthreads();
        QR_STAMP(15);
        // SA <- M = U R2 (upper triangular), SB <- M^-1: rows below the top block are  Q1 M^-1
        {
          T out[4];
#pragma unroll
          for (int kk = 0; kk < 4; ++kk) {
            const int k = k0 + kk;
            T sum = N_::zero();
            for (int q = i; q <= k; ++q) sum = N_::fma(SC[i * QP + q], SB[q * QP + k], sum);
            out[kk] = sum;
          }
          __syncthreads();
#pragma unroll
          for (int kk = 0; kk < 4; ++kk) SA[i * QP + k0 + kk] = (k0 + kk >= i) ? out[kk] : N_::zero();
          __syncthreads();
        }
        panel_tri_inverse<T>(SA, SB, SC, tid);
        QR_STAMP(16);
      }
      cluster.sync();  // CTA 0's SB holds (U R2)^-1
      if (crank != 0) {
        const T* SB0 = cluster.map_shared_rank(SB, 0);
        for (int idx = tid; idx < QR_NB * QR_NB; idx += QR_THREADS) {
          const int r = idx / QR_NB, k = idx - r * QR_NB;
          SA[r * QP + k] = SB0[r * QP + k];
        }
        __syncthreads();
      }
      cluster.sync();  // every CTA holds its own copy: CTA 0 may finish and exit whenever it likes
      QR_STAMP(17);
      panel_right_mult<T>(P, a.pitch, crank == 0 ? QR_NB : 0, nrows, crank == 0 ? SB : SA, warp, lane);
      __syncthreads();
      QR_STAMP(18);
      for (int idx = tid; idx < nrows * jb; idx += QR_THREADS) {
        const int r = idx / jb, c = idx - r * jb;
        Vg[(a.j0 + r_lo + r) * a.ldv + a.j0 + c] = P[c * a.pitch + r];
      }
      QR_STAMP(19);
      return;
    }
    // ill-conditioned panel: restore the slab and fall through to the Householder loop
    __syncthreads();
    for (int i = tid; i < nrows * jb; i += QR_THREADS) {
      const int r = i / jb, c = i - r * jb;
      P[c * a.pitch + r] = Wg[(a.j0 + r_lo + r) * a.ldw + a.j0 + c];
    }
    __syncthreads();
  }

  // stage[r][c]: all CTAs' partial dots of the current column step, copied from DSMEM in ONE round trip
  T* stage = Z;  // QR_NB x QR_NB scratch: Z is only needed after the column loop (C <= 16 ranks <= QR_NB rows)
  __shared__ T rowv_l[QR_NB];
  for (int j = 0; j < jb; ++j) {
    const int buf = j & 1;
    // local index of the first row at or below the diagonal
    int lo = (int)(j - r_lo);
    if (lo < 0) lo = 0;
    const bool own_diag = (j >= r_lo && j < r_lo + nrows);
    const int jl = (int)(j - r_lo);  // local row of the diagonal (valid if own_diag)
    // ---- partial dots d_c = sum_{r >= j} conj(x_r) A[r, c], c = j..jb-1: each warp takes columns
    //      c = j + warp + 8 i (up to 4) and reduces them together ------------------------------------
    {
      const T* xj = P + j * a.pitch;
      T acc[4];
      const T* xc[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        acc[i] = N_::zero();
        const int c = j + warp + QR_WARPS * i;
        xc[i] = P + (c < jb ? c : j) * a.pitch;
      }
#pragma unroll 4
      for (int r = lo + lane; r < nrows; r += 32) {
        const T x = xj[r];
#pragma unroll
        for (int i = 0; i < 4; ++i) acc[i] = N_::fma_conj(x, xc[i][r], acc[i]);
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          if constexpr (sizeof(T) == 16) {
            acc[i].x += __shfl_xor_sync(0xffffffffu, acc[i].x, o);
            acc[i].y += __shfl_xor_sync(0xffffffffu, acc[i].y, o);
          } else {
            acc[i] += __shfl_xor_sync(0xffffffffu, acc[i], o);
          }
        }
      }
      if (lane == 0) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int c = j + warp + QR_WARPS * i;
          if (c < jb) {
            part[buf * QR_NB + c] = acc[i];
            if (own_diag) rowv[buf * QR_NB + c] = xc[i][jl];
          }
        }
      }
    }
    cluster.sync();
    // ---- one DSMEM round trip: every CTA copies all partials (and the diagonal row) locally ------
    const int owner = (int)(j / a.rows_per);
    for (int t = tid; t < C * QR_NB; t += QR_THREADS) {
      const int rk = t / QR_NB, c = t - rk * QR_NB;
      if (c >= j && c < jb) stage[rk * QR_NB + c] = cluster.map_shared_rank(part, rk)[buf * QR_NB + c];
    }
    if (tid >= QR_THREADS - QR_NB) {
      const int c = tid - (QR_THREADS - QR_NB);
      if (c >= j && c < jb) rowv_l[c] = cluster.map_shared_rank(rowv, owner)[buf * QR_NB + c];
    }
    __syncthreads();
    // ---- every thread forms the same reflector parameters ---------------------------------------
    T dj = N_::zero();
    for (int rk = 0; rk < C; ++rk) dj = N_::add(dj, stage[rk * QR_NB + j]);
    const T alpha = rowv_l[j];
    const double normx2 = N_::real(dj);
    const bool no_tail = (a.j0 + j == a.m - 1);
    double beta;
    T tau_j, scale;  // v = x * scale below the diagonal
    double alpha_im = 0.0;
    if constexpr (sizeof(T) == 16) alpha_im = alpha.y;
    if (normx2 == 0.0 || (no_tail && alpha_im == 0.0)) {
      beta = N_::real(alpha);
      tau_j = N_::zero();
      scale = N_::zero();
    } else {
      beta = -copysign(sqrt(normx2), N_::real(alpha));
      const double ib = 1.0 / beta;
      tau_j = N_::from((beta - N_::real(alpha)) * ib, -alpha_im * ib);
      // scale = 1 / (alpha - beta)
      const double dr = N_::real(alpha) - beta;
      const double den = dr * dr + alpha_im * alpha_im;
      scale = N_::from(dr / den, -alpha_im / den);
    }
    if (tid == 0) tau[j] = tau_j;
    const bool active = !(N_::real(tau_j) == 0.0 && N_::abs2(tau_j) == 0.0);
    const int lo1 = own_diag ? jl + 1 : lo;
    // ---- C2: A[:, c] -= conj(tau) * (v^H A[:, c]) * v, c > j, with v = x_j * scale read on the fly ----
    if (active) {
      const T* xj = P + j * a.pitch;
      for (int c = j + 1 + warp; c < jb; c += QR_WARPS) {
        T dc = N_::zero();
        for (int rk = 0; rk < C; ++rk) dc = N_::add(dc, stage[rk * QR_NB + c]);
        const T ajc = rowv_l[c];
        // w = v^H A_c = A[j,c] + conj(scale) * (d_c - conj(alpha) * A[j,c])
        const T t1 = N_::sub(dc, N_::mul(N_::conj(alpha), ajc));
        const T w = N_::add(ajc, N_::mul(N_::conj(scale), t1));
        const T f = N_::mul(N_::conj(tau_j), w);
        const T fs = N_::mul(f, scale);
        T* xcc = P + c * a.pitch;
#pragma unroll 4
        for (int r = lo1 + lane; r < nrows; r += 32) xcc[r] = N_::sub(xcc[r], N_::mul(fs, xj[r]));
        if (own_diag && lane == 0) xcc[jl] = N_::sub(xcc[jl], f);  // v_j = 1 on the diagonal
      }
    }
    __syncthreads();
    // ---- C1: x -> v on column j, diagonal := beta (column j is not read again by later columns) ------
    {
      T* xj = P + j * a.pitch;
      for (int r = lo1 + tid; r < nrows; r += QR_THREADS) xj[r] = N_::mul(xj[r], scale);
      if (own_diag && tid == 0) xj[jl] = N_::from(beta, 0.0);
    }
  }
  __syncthreads();

  // ---- R goes out; P becomes the explicit V (zeros above the diagonal, ones on it, zero pad rows) ----
  const int K4 = (nrows + 3) & ~3;
  for (int i = tid; i < nrows * jb; i += QR_THREADS) {
    const int r = i / jb, c = i - r * jb;
    const int64_t pr = r_lo + r;  // panel row
    if (pr <= c) {
      Wg[(a.j0 + pr) * a.ldw + a.j0 + c] = P[c * a.pitch + r];
      P[c * a.pitch + r] = (pr == c) ? N_::one() : N_::zero();
    }
  }
  for (int i = tid; i < (K4 - nrows) * QR_NB; i += QR_THREADS) {
    const int c = i / (K4 - nrows), r = nrows + (i - c * (K4 - nrows));
    P[c * a.pitch + r] = N_::zero();
  }
  __syncthreads();
  // ---- T factor.  Z = V^H V over this CTA's rows on the FP64 tensor pipe (4 x 4 tiles of 8 x 8, upper
  //      tiles only, two per warp), summed over the cluster; T = (striu(Z) + diag(1/tau))^-1. ---------
  {
    constexpr bool CPLX = (sizeof(T) == 16);
    const int gq = lane >> 2, tq = lane & 3;
    const int gm = warp >> 1, gn0 = (warp & 1) * 2;
    double g[2][CPLX ? 4 : 2];
#pragma unroll
    for (int j = 0; j < 2; ++j)
#pragma unroll
      for (int r = 0; r < (CPLX ? 4 : 2); ++r) g[j][r] = 0.0;
    if (gn0 + 1 >= gm) {
      const T* pa = P + (gm * 8 + gq) * a.pitch + tq;
      const T* pb0 = P + (gn0 * 8 + gq) * a.pitch + tq;
      const T* pb1 = pb0 + 8 * a.pitch;
#pragma unroll 4
      for (int k0 = 0; k0 < K4; k0 += 4) {
        const T av = pa[k0], b0 = pb0[k0], b1 = pb1[k0];
        if constexpr (CPLX) {
          const double nay = -av.y;
          dmma884(g[0][0], g[0][1], av.x, b0.x);
          dmma884(g[0][2], g[0][3], av.x, b0.y);
          dmma884(g[1][0], g[1][1], av.x, b1.x);
          dmma884(g[1][2], g[1][3], av.x, b1.y);
          dmma884(g[0][0], g[0][1], av.y, b0.y);
          dmma884(g[0][2], g[0][3], nay, b0.x);
          dmma884(g[1][0], g[1][1], av.y, b1.y);
          dmma884(g[1][2], g[1][3], nay, b1.x);
        } else {
          dmma884(g[0][0], g[0][1], av, b0);
          dmma884(g[1][0], g[1][1], av, b1);
        }
      }
    }
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      const int r = gm * 8 + gq, c = (gn0 + j) * 8 + 2 * tq;
      if constexpr (CPLX) {
        Z[r * QR_NB + c] = make_double2(g[j][0], g[j][2]);
        Z[r * QR_NB + c + 1] = make_double2(g[j][1], g[j][3]);
      } else {
        Z[r * QR_NB + c] = g[j][0];
        Z[r * QR_NB + c + 1] = g[j][1];
      }
    }
  }
  cluster.sync();
  if (crank == 0) {
    for (int idx = tid; idx < jb * jb; idx += QR_THREADS) {
      const int i = idx / jb, c = idx - i * jb;
      if (i >= c) continue;
      T sum = N_::zero();
      for (int q = 0; q < C; ++q) sum = N_::add(sum, cluster.map_shared_rank(Z, q)[i * QR_NB + c]);
      Z[i * QR_NB + c] = sum;  // only CTA 0's copy holds the total; the others are just read
    }
  }
  cluster.sync();  // totals complete, and no CTA exits while its Z is still being read
  if (crank == 0 && warp == 0) {
    // T is upper triangular with T^-1 = striu(V^H V) + diag(1/tau) (equivalent to LAPACK larft, forward /
    // columnwise).  Lane c solves for column c by back substitution:
    //   t_c = tau_c,  t_i = -tau_i * sum_{q = i+1..c} z[i][q] t_q   (tau_i = 0 gives a zero row, as larft).
    // Column c is kept in the lower triangle of Z, transposed: T[i][c] lives at Z[c][i], i <= c, so the lanes
    // never touch each other's entries nor the strictly upper part that holds z.
    const int c = lane;
    if (c < jb) {
      Z[c * QR_NB + c] = tau[c];
      for (int i = c - 1; i >= 0; --i) {
        T sum = N_::zero();
        for (int q = i + 1; q <= c; ++q) sum = N_::fma(Z[i * QR_NB + q], Z[c * QR_NB + q], sum);
        Z[c * QR_NB + i] = N_::mul(N_::sub(N_::zero(), tau[i]), sum);
      }
    }
    __syncwarp();
    T* Tg = reinterpret_cast<T*>(a.T);
    for (int idx = lane; idx < jb * jb; idx += 32) {
      const int i = idx / jb, cc = idx - i * jb;
      T v = N_::zero();
      if (i <= cc) v = Z[cc * QR_NB + i];
      Tg[i * a.ldt + cc] = v;
    }
  }

  // ---- explicit V for every row ------------------------------------------------------------------
  for (int i = tid; i < nrows * jb; i += QR_THREADS) {
    const int r = i / jb, c = i - r * jb;
    Vg[(a.j0 + r_lo + r) * a.ldv + a.j0 + c] = P[c * a.pitch + r];
  }
}

// max |re|, |im| over a matrix -> atomicMax on the bit pattern (non-negative doubles order like integers)
template <typename T>
__global__ void amax_kernel(const T* src, int64_t lds, int64_t rows, int64_t cols, unsigned long long* out) {
  const int64_t step = (int64_t)gridDim.x * blockDim.x;
  double m = 0.0;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < rows * cols; i += step) {
    const int64_t r = i / cols, c = i - r * cols;
    const T v = src[r * lds + c];
    double a;
    if constexpr (sizeof(T) == 16) a = fmax(fabs(v.x), fabs(v.y)); else a = fabs(v);
    if (a > m) m = a;  // NaN never wins; it propagates through the factorisation itself
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const double other = __shfl_xor_sync(0xffffffffu, m, o);
    if (other > m) m = other;
  }
  if ((threadIdx.x & 31) == 0 && m > 0.0) atomicMax(out, (unsigned long long)__double_as_longlong(m));
}
// scale[0] = 2^-ilogb(amax) (exact power of two: scaling by it is rounding-free), 1 for zero / non-finite
__global__ void pow2_scale_kernel(const unsigned long long* amax_bits, double* scale) {
  const double a = __longlong_as_double((long long)amax_bits[0]);
  double s = 1.0;
  if (a > 0.0 && isfinite(a)) s = scalbn(1.0, -ilogb(a));
  scale[0] = s;
  scale[1] = 1.0 / s;
}
// copy A (lda) -> W (ldw), contiguous rows, times scale[0] when given
template <typename T>
__global__ void copy2d_kernel(const T* src, int64_t lds, T* dst, int64_t ldd, int64_t rows, int64_t cols,
                              const double* scale, const int* skip = nullptr) {
  griddep_wait();               // no-ops unless launched with programmatic stream serialization (launch_k)
  griddep_launch_dependents();
  if (skip && *skip) return;    // device-side skip (SkipScope)
  const int64_t step = (int64_t)gridDim.x * blockDim.x;
  const double s = scale ? scale[0] : 1.0;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < rows * cols; i += step) {
    const int64_t r = i / cols, c = i - r * cols;
    dst[r * ldd + c] = Num<T>::scale(src[r * lds + c], s);
  }
}
// W (m x b, ld ldw) <- W * Rinv (b x b upper triangular, ld ldr), in place, b <= 64: the last step of a first pass
// of qr_bcgs2.  As a GEMM this is m x 64 x 64 -- eight k-steps, so pipeline fill, epilogue and the copy back from a
// second buffer cost more than the product (19 + 4 us measured).  Here one CTA of eight warps owns 32 rows, staged
// once in shared memory (pitch as the GEMM's K-contiguous tile: conflict-free DMMA fragments); warp j forms the
// eight columns 8j .. 8j + 7 of all 32 rows on DMMA and keeps ITS columns of Rinv -- only the k <= 8j + 7 part of
// the triangular factor -- as B fragments in registers, loaded once from global memory; the results go back over
// the CTA's own rows.
constexpr int PS_ROWS = 32, PS_N = 64, PS_WP = PS_N + 4;
template <typename T>
__global__ void __launch_bounds__(256) panel_scale_kernel(T* W, int64_t ldw, int64_t m, int b, const T* Rinv, int64_t ldr) {
  typedef Num<T> N_;
  constexpr bool CPLX = sizeof(T) == 16;
  __shared__ __align__(16) T Ws[PS_ROWS * PS_WP];   // rows of W, k contiguous
  griddep_wait();
  griddep_launch_dependents();
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, gq = lane >> 2, tq = lane & 3;
  const int64_t r0 = (int64_t)blockIdx.x * PS_ROWS;
  // B fragments of this warp's column tile: Rinv[k = k0 + tq][n = 8 warp + gq], zero outside the triangle / b
  T bf[PS_N / 4];
#pragma unroll
  for (int ks = 0; ks < PS_N / 4; ++ks) {
    const int kk = ks * 4 + tq, n = warp * 8 + gq;
    bf[ks] = (ks * 4 <= warp * 8 + 7 && kk <= n && n < b) ? Rinv[(int64_t)kk * ldr + n] : N_::zero();
  }
  for (int idx = tid; idx < PS_ROWS * PS_N; idx += 256) {
    const int r = idx / PS_N, kk = idx - r * PS_N;
    Ws[r * PS_WP + kk] = (r0 + r < m && kk < b) ? W[(r0 + r) * ldw + kk] : N_::zero();
  }
  __syncthreads();
  constexpr int NACC = CPLX ? 6 : 2;
  double acc[PS_ROWS / 8][NACC];
#pragma unroll
  for (int i = 0; i < PS_ROWS / 8; ++i)
#pragma unroll
    for (int r = 0; r < NACC; ++r) acc[i][r] = 0.0;
#pragma unroll
  for (int ks = 0; ks < PS_N / 4; ++ks) {
    if (ks * 4 > warp * 8 + 7) break;        // Rinv[k][n] = 0 for k > n (uniform over the warp)
    const T bv = bf[ks];
    double bs = 0.0;
    if constexpr (CPLX) bs = bv.x + bv.y;
#pragma unroll
    for (int i = 0; i < PS_ROWS / 8; ++i) {
      const T av = Ws[(i * 8 + gq) * PS_WP + ks * 4 + tq];
      if constexpr (CPLX) {
        dmma884(acc[i][0], acc[i][1], av.x, bv.x);
        dmma884(acc[i][2], acc[i][3], av.y, bv.y);
        dmma884(acc[i][4], acc[i][5], av.x + av.y, bs);
      } else {
        dmma884(acc[i][0], acc[i][1], av, bv);
      }
    }
  }
  const int col = warp * 8 + 2 * tq;
#pragma unroll
  for (int i = 0; i < PS_ROWS / 8; ++i) {
    const int64_t row = r0 + i * 8 + gq;
    if (row >= m) continue;
    T v0, v1;
    if constexpr (CPLX) {
      v0 = make_double2(acc[i][0] - acc[i][2], acc[i][4] - acc[i][0] - acc[i][2]);
      v1 = make_double2(acc[i][1] - acc[i][3], acc[i][5] - acc[i][1] - acc[i][3]);
    } else {
      v0 = acc[i][0]; v1 = acc[i][1];
    }
    if (col < b) W[row * ldw + col] = v0;
    if (col + 1 < b) W[row * ldw + col + 1] = v1;
  }
}
// R = triu(W[0:k, 0:n]) (times unscale[1] when given)
template <typename T>
__global__ void extract_r_kernel(const T* W, int64_t ldw, T* R, int64_t k, int64_t n, const double* unscale) {
  const int64_t step = (int64_t)gridDim.x * blockDim.x;
  const double s = unscale ? unscale[1] : 1.0;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < k * n; i += step) {
    const int64_t r = i / n, c = i - r * n;
    R[i] = (c >= r) ? Num<T>::scale(W[r * ldw + c], s) : Num<T>::zero();
  }
}
// Q := [I_k; 0] (m x k, ld)
template <typename T>
__global__ void eye_kernel(T* Q, int64_t m, int64_t k, int64_t ld) {
  const int64_t step = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < m * k; i += step) {
    const int64_t r = i / k, c = i - r * k;
    Q[r * ld + c] = (r == c) ? Num<T>::one() : Num<T>::zero();
  }
}

static inline unsigned blocks_for(int64_t n) {
  int64_t b = (n + 255) / 256;
  const int64_t cap = (int64_t)sm_count() * 16;
  if (b > cap) b = cap;
  if (b < 1) b = 1;
  return (unsigned)b;
}

static inline size_t align_up(size_t x) { return (x + 255) & ~(size_t)255; }

// Compile-time switches of the kernel experiments (scratch/): the release build has no run-time knobs.
#ifndef TNB_EXP_QR_FAST_PANEL
#define TNB_EXP_QR_FAST_PANEL 1   // 0: column-by-column Householder panel everywhere
#endif
#ifndef TNB_EXP_QR_OVERLAP
#define TNB_EXP_QR_OVERLAP 1      // 0: everything on the caller's stream
#endif
#ifndef TNB_EXP_QR_SIDE_PCT
#define TNB_EXP_QR_SIDE_PCT 20    // share of the split-K scratch given to the side stream
#endif
#ifndef TNB_EXP_QR_BCGS
#define TNB_EXP_QR_BCGS 2         // 0: Householder path everywhere; 1: no lagged second pass
#endif
constexpr int g_qr_fast_panel = TNB_EXP_QR_FAST_PANEL;

constexpr int QR_NBO = 128;  // outer block: trailing updates and the explicit Q use K = 128 GEMMs

struct QrLayout {
  int64_t k, ldw, ldv, nouter;
  size_t off_w, off_v, off_t, off_w1, off_w2, off_sc, off_sk, sk_bytes, off_p2, off_cb, off_bc, off_rt, total;
};

constexpr int QR_CB = 64;   // column block of the Gram-Schmidt path (qr_bcgs2)
constexpr int QR_GB = 256;  // column group of its lagged second pass; also the pitch of its panel buffers
// A group whose Gram matrix against [earlier groups, itself] is within this of I after the FIRST pass is left as it is
// (the update kernels of its second pass return at once): 1e-14 is the level of LAPACK's own Householder Q for these
// sizes (eps sqrt(m) = 1.2e-14 at m = 3072).  Measured on the cfg 3 sweep: two groups in three qualify.
#ifndef TNB_EXP_QR_SKIP_TOL
#define TNB_EXP_QR_SKIP_TOL 1e-14
#endif
constexpr double QR_SKIP_TOL = TNB_EXP_QR_SKIP_TOL;

static QrLayout qr_layout(int dtype, int64_t m, int64_t n) {
  QrLayout L;
  const size_t es = elem_size(dtype);
  L.k = m < n ? m : n;
  L.ldw = (n + 1) & ~(int64_t)1;
  L.ldv = (L.k + 1) & ~(int64_t)1;
  L.nouter = (L.k + QR_NBO - 1) / QR_NBO;
  const int64_t wide = (n > L.k ? n : L.k);
  size_t o = 0;
  L.off_w = o;  o += align_up((size_t)m * L.ldw * es);
  L.off_v = o;  o += align_up((size_t)m * L.ldv * es);
  L.off_t = o;  o += align_up((size_t)L.nouter * QR_NBO * QR_NBO * es);
  L.off_w1 = o; o += align_up((size_t)QR_NBO * (wide + 2) * es);
  L.off_w2 = o; o += align_up((size_t)QR_NBO * (wide + 2) * es);
  L.off_sc = o; o += 256;  // [0] amax bits, [1] scale, [2] 1/scale
  // split-K scratch for the (jb x n2 x m) products: up to 16 partials of a 128 x wide block
  L.sk_bytes = (m >= 1024) ? align_up((size_t)16 * QR_NBO * wide * es) : 0;
  L.off_sk = o; o += L.sk_bytes;
  L.off_p2 = o; o += align_up((size_t)m * QR_GB * es);    // qr_bcgs2: second panel buffer
  L.off_cb = o; o += align_up((size_t)L.k * QR_GB * es);  // qr_bcgs2: [Q^H W; W^H W] of the second pass
  L.off_bc = o; o += align_up((size_t)L.k * QR_GB * es);  // qr_bcgs2: [-C R^-1; R^-1] of the current pass
  L.off_rt = o; o += align_up((size_t)3 * QR_GB * QR_GB * es);  // qr_bcgs2: small factors
  L.total = o;
  return L;
}

template <typename T>
static int launch_panel(QrPanelArgs& a, cudaStream_t st) {
  const int64_t mr = a.m - a.j0;
  // smallest cluster whose slabs fit in shared memory (cap 200 KB per CTA)
  const size_t fixed = (size_t)(2 * QR_NB + 2 * QR_NB + QR_NB + QR_NB * QR_NB + 3 * QR_NB * QP) * sizeof(T);
  int C = 1;
  int rows_per = 0, pitch = 0;
  size_t smem = 0;
  for (;; C *= 2) {
    rows_per = (int)((mr + C - 1) / C);
    pitch = ((rows_per + 3) & ~3) + 1;  // odd (conflict-free transposing load), with room for the k4 zero padding
    smem = (size_t)QR_NB * pitch * sizeof(T) + fixed;
    if (smem <= 220 * 1024) break;
    if (C >= 16) return TNB_E_UNSUPPORTED;  // panel taller than 16 CTAs can hold
  }
  a.rows_per = rows_per;
  a.pitch = pitch;
  auto kern = qr_panel_kernel<T>;
  static PerDeviceOnce once;
  if (once.need()) {
    TNB_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(220 * 1024)));
    TNB_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
    once.done();
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)C, 1, 1);
  cfg.blockDim = dim3(QR_THREADS, 1, 1);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = (unsigned)C;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  // panel factorisation + T factor: ~ (2 jb^2 + jb^2) mr real flops, x4 complex
  ProfScope prof(KC_QR_PANEL, st, (sizeof(T) == 16 ? 4.0 : 1.0) * 3.0 * (double)mr * a.jb * a.jb);
  TNB_CUDA_CHECK(cudaLaunchKernelEx(&cfg, kern, a));
  count_launch();
  return 0;
}

// ---------------------------------------------------------------------------------------------------
// Block classical Gram-Schmidt with reorthogonalisation and Pythagorean inner products (BCGS-PIP+, Carson,
// Lund, Rozloznik, Thomas 2021) -- the QR path for matrices with many columns.  Every O(m n^2) operation
// is a DMMA GEMM over all SMs; the only small serial kernel is the 64 x 64 Cholesky factor + triangular
// inverse below (two per column block).  With Qj = Q[:, :j] and W = A[:, J] stored next to it as Q[:, J],
// one pass over a block J of QR_CB columns is two large GEMMs:
//     [C; G0] = [Qj, W]^H W              (K = m, split-K)
//     G  = G0 - C^H C = (W - Qj C)^H (W - Qj C) = R^H R            (Pythagoras; 64 x 64 kernel)
//     W <- (W - Qj C) R^-1 = [Qj, W] [-C R^-1; R^-1]               (K = j + 64)
// and two passes give  Q[:, J] = W,  R[J, J] = R2 R1,  R[:j, J] = C1 + C2 R1.
// A first-pass Cholesky pivot below 1e-10 of its diagonal entry, or a second-pass Gram matrix that is not
// close to I (pivot < 1/4: the first pass lost orthogonality -- cancellation in the Pythagorean step,
// numerically dependent columns), raises a device flag; the caller then runs the Householder path on the
// untouched A.
// ---------------------------------------------------------------------------------------------------
template <typename T> struct CPLX_CI { static constexpr bool value = sizeof(T) == 16; };
constexpr int CI_N = QR_CB;       // matrix order handled by chol_inv_kernel
constexpr int CI_P = CI_N + 1;    // shared-memory pitch
constexpr int CI_THREADS = 512;

// G (n x n Hermitian, n <= 64; upper triangle read) -> R (upper, G = R^H R, real positive diagonal) and
// Rinv = R^-1 (upper); strictly lower triangles are written as zeros.  One CTA.  flag[0] is set when a pivot
// is <= rel_floor * (its original diagonal entry) or <= abs_floor.  With near_tol > 0, a G within near_tol of
// the identity (entrywise) takes the first-order formulas instead of the 64 pivot steps.
template <typename T>
__global__ void __launch_bounds__(CI_THREADS) chol_inv_kernel(const T* G, int64_t ldg, int n, T* R, int64_t ldr,
                                                              T* Rinv, int64_t ldri, double rel_floor,
                                                              double abs_floor, double near_tol, int* flag) {
  typedef Num<T> N_;
  extern __shared__ __align__(16) unsigned char ci_smem[];
  T* A = reinterpret_cast<T*>(ci_smem);  // [CI_N][CI_P]  G -> R
  T* X = A + CI_N * CI_P;                // [CI_N][CI_P]  R^-1
  T* Y = X + CI_N * CI_P;                // [CI_N][CI_P]  scratch of the inversion
  __shared__ double diag0[CI_N];
  __shared__ int s_bad;
  const int tid = threadIdx.x;
  if (tid == 0) s_bad = 0;
  griddep_wait();
  griddep_launch_dependents();
  for (int idx = tid; idx < CI_N * CI_N; idx += CI_THREADS) {
    const int i = idx / CI_N, j = idx - i * CI_N;
    T v = N_::zero();
    if (i < n && j < n) { if (j >= i) v = G[(int64_t)i * ldg + j]; }
    else if (i == j) v = N_::one();  // identity padding
    A[i * CI_P + j] = v;
  }
  // Second-pass Gram matrices are I + E with |E| ~ eps * cond(first-pass block)^2, usually ~1e-13: then
  // R = I + U and R^-1 = I - U with U = striu(E) + diag(E)/2 are exact to O(|E|^2) -- no pivot steps at all.
  {
    int slow = (near_tol > 0.0) ? 0 : 1;
    if (!slow) {
      for (int idx = tid; idx < CI_N * CI_N; idx += CI_THREADS) {
        const int i = idx / CI_N, j = idx - i * CI_N;
        if (j < i) continue;
        const T d = N_::sub(A[i * CI_P + j], (i == j) ? N_::one() : N_::zero());
        if (!(N_::abs2(d) <= near_tol * near_tol)) slow = 1;  // also true for NaN
      }
    }
    if (!__syncthreads_or(slow)) {
      for (int idx = tid; idx < n * n; idx += CI_THREADS) {
        const int i = idx / n, j = idx - i * n;
        T r = N_::zero(), ri = N_::zero();
        if (j > i) { r = A[i * CI_P + j]; ri = N_::sub(N_::zero(), r); }
        else if (j == i) {
          const double h = 0.5 * (N_::real(A[i * CI_P + i]) - 1.0);
          r = N_::from(1.0 + h, 0.0);
          ri = N_::from(1.0 - h, 0.0);
        }
        R[(int64_t)i * ldr + j] = r;
        Rinv[(int64_t)i * ldri + j] = ri;
      }
      return;
    }
  }
  if (tid < CI_N) diag0[tid] = N_::real(A[tid * CI_P + tid]);
  // Right-looking Cholesky on the upper triangle, rows kept unscaled (Schur complements: U, with R = D^-1/2 U),
  // blocked by 8 rows.  The 8 pivot steps of a block touch only its own 8 rows (thread = (row r of the block,
  // column c): one complex multiply-add per step, so a step is a short dependent chain instead of 32 FP64
  // multiply-adds per thread); the rank-8 update of everything below is then done on DMMA by all warps:
  //   A[i][k] -= sum_p conj(U[p][i]) / d_p * U[p][k]      (i <= k: upper 8 x 8 tiles only)
  const int k = tid & (CI_N - 1), i0 = (tid / CI_N) * (CI_N * CI_N / CI_THREADS);
  constexpr int RPT = CI_N * CI_N / CI_THREADS;  // rows per thread of the final scaling (8)
  __shared__ double s_ipiv[CI_N];
  const int warp = tid >> 5, lane = tid & 31, gq = lane >> 2, tq = lane & 3;
  bool bad = false;
  for (int kb = 0; kb < CI_N / 8; ++kb) {
    const int j0 = kb * 8;
    const int pr = tid >> 6, pc = tid & 63;   // panel row (0..7) and column of this thread
    for (int t = 0; t < 8; ++t) {
      const int j = j0 + t;
      __syncthreads();
      const double piv = N_::real(A[j * CI_P + j]);
      if (!(piv > rel_floor * diag0[j]) || !(piv > abs_floor)) bad = true;
      const double ipiv = __drcp_rn(piv);
      if (tid == 0) s_ipiv[j] = ipiv;
      const int i = j0 + pr;
      if (pr > t && pc >= i)
        A[i * CI_P + pc] = N_::sub(A[i * CI_P + pc], N_::mul(N_::conj(A[j * CI_P + i]), N_::scale(A[j * CI_P + pc], ipiv)));
    }
    __syncthreads();
    const int nt = CI_N / 8 - 1 - kb;   // tiles per dimension of the trailing matrix
    for (int tile = warp; tile < nt * (nt + 1) / 2; tile += CI_THREADS / 32) {
      int mi = 0, rem = tile;           // tile (mi <= nj) of the upper triangle, row-major
      while (rem >= nt - mi) { rem -= nt - mi; ++mi; }
      const int nj = mi + rem;
      const int I0 = j0 + 8 + mi * 8, K0 = j0 + 8 + nj * 8;
      double acc[CPLX_CI<T>::value ? 4 : 2] = {};
#pragma unroll
      for (int p0 = 0; p0 < 8; p0 += 4) {
        const int p = j0 + p0 + tq;
        const T a = N_::scale(N_::conj(A[p * CI_P + I0 + gq]), s_ipiv[p]);
        const T b = A[p * CI_P + K0 + gq];
        if constexpr (CPLX_CI<T>::value) {
          dmma884(acc[0], acc[1], a.x, b.x);
          dmma884(acc[0], acc[1], -a.y, b.y);
          dmma884(acc[2], acc[3], a.x, b.y);
          dmma884(acc[2], acc[3], a.y, b.x);
        } else {
          dmma884(acc[0], acc[1], a, b);
        }
      }
      T* c = A + (I0 + gq) * CI_P + K0 + 2 * tq;
      if constexpr (CPLX_CI<T>::value) {
        c[0] = N_::sub(c[0], make_double2(acc[0], acc[2]));
        c[1] = N_::sub(c[1], make_double2(acc[1], acc[3]));
      } else {
        c[0] -= acc[0];
        c[1] -= acc[1];
      }
    }
  }
  __syncthreads();
  if (bad && tid == 0) s_bad = 1;
  double rs[RPT];
#pragma unroll
  for (int ii = 0; ii < RPT; ++ii) rs[ii] = rsqrt(N_::real(A[(i0 + ii) * CI_P + i0 + ii]));
  __syncthreads();
#pragma unroll
  for (int ii = 0; ii < RPT; ++ii) {
    const int i = i0 + ii;
    T v = (k >= i) ? N_::scale(A[i * CI_P + k], rs[ii]) : N_::zero();
    if (k == i) v = N_::from(N_::real(v), 0.0);
    A[i * CI_P + k] = v;
    // inverse of the diagonal: 1 / R[i][i] = rs * (1 / (d rs^2)) ... R[i][i] = d * rs with d the pivot: 1/R = 1/(d rs)
    X[i * CI_P + k] = (k == i) ? N_::from(1.0 / N_::real(v), 0.0) : N_::zero();
  }
  __syncthreads();
  if (s_bad) {
    if (tid == 0) flag[0] = 1;
    // still write finite output (identity) so that the queued GEMMs of the caller stay harmless
    for (int idx = tid; idx < n * n; idx += CI_THREADS) {
      const int i = idx / n, j = idx - i * n;
      R[(int64_t)i * ldr + j] = (i == j) ? N_::one() : N_::zero();
      Rinv[(int64_t)i * ldri + j] = (i == j) ? N_::one() : N_::zero();
    }
    return;
  }
  // X = A^-1 by recursive doubling over diagonal blocks of size b = 1, 2, .. 32:
  //   X12 = -X11 (A12 X22) for every pair of adjacent diagonal blocks
  // b < 8: one thread per entry; b >= 8: the two products run on DMMA, one 8 x 8 output tile per warp
  // (32 / b block pairs x (b / 8)^2 tiles = b / 2 tiles; the zeros below the diagonals of X11 / X22 are stored)
  for (int b = 1; b < 8; b *= 2) {
    const int nent = CI_N * b / 2;
    for (int e = tid; e < nent; e += CI_THREADS) {
      const int blk = e / (b * b), r = (e / b) % b, c = e % b;
      const int s0 = blk * 2 * b;
      T sum = N_::zero();
      for (int q = 0; q <= c; ++q) sum = N_::fma(A[(s0 + r) * CI_P + s0 + b + q], X[(s0 + b + q) * CI_P + s0 + b + c], sum);
      Y[(s0 + r) * CI_P + s0 + b + c] = sum;
    }
    __syncthreads();
    for (int e = tid; e < nent; e += CI_THREADS) {
      const int blk = e / (b * b), r = (e / b) % b, c = e % b;
      const int s0 = blk * 2 * b;
      T sum = N_::zero();
      for (int q = r; q < b; ++q) sum = N_::fma(X[(s0 + r) * CI_P + s0 + q], Y[(s0 + q) * CI_P + s0 + b + c], sum);
      X[(s0 + r) * CI_P + s0 + b + c] = N_::sub(N_::zero(), sum);
    }
    __syncthreads();
  }
  // C[rc.., cc..] = sgn * P[rp.., cp..] (b x b) * Q[rq.., cq..] (b x b), tile (mi, nj) of 8 x 8, all in shared memory
  auto tile_product = [&](T* Cm, int rc, int cc, const T* Pm, int rp, int cp, const T* Qm, int rq, int cq, int b, int mi,
                          int nj, double sgn) {
    double acc[CPLX_CI<T>::value ? 4 : 2] = {};
    for (int k0 = 0; k0 < b; k0 += 4) {
      const T a = Pm[(rp + mi * 8 + gq) * CI_P + cp + k0 + tq];
      const T q = Qm[(rq + k0 + tq) * CI_P + cq + nj * 8 + gq];
      if constexpr (CPLX_CI<T>::value) {
        dmma884(acc[0], acc[1], a.x, q.x);
        dmma884(acc[0], acc[1], -a.y, q.y);
        dmma884(acc[2], acc[3], a.x, q.y);
        dmma884(acc[2], acc[3], a.y, q.x);
      } else {
        dmma884(acc[0], acc[1], a, q);
      }
    }
    T* c = Cm + (rc + mi * 8 + gq) * CI_P + cc + nj * 8 + 2 * tq;
    if constexpr (CPLX_CI<T>::value) {
      c[0] = make_double2(sgn * acc[0], sgn * acc[2]);
      c[1] = make_double2(sgn * acc[1], sgn * acc[3]);
    } else {
      c[0] = sgn * acc[0];
      c[1] = sgn * acc[1];
    }
  };
  for (int b = 8; b < CI_N; b *= 2) {
    const int tpb = (b / 8) * (b / 8), ntile = (CI_N / (2 * b)) * tpb;
    for (int tile = warp; tile < ntile; tile += CI_THREADS / 32) {
      const int blk = tile / tpb, tt = tile - blk * tpb, mi = tt / (b / 8), nj = tt - mi * (b / 8);
      const int s0 = blk * 2 * b;
      tile_product(Y, s0, s0 + b, A, s0, s0 + b, X, s0 + b, s0 + b, b, mi, nj, 1.0);    // Y12 = A12 X22
    }
    __syncthreads();
    for (int tile = warp; tile < ntile; tile += CI_THREADS / 32) {
      const int blk = tile / tpb, tt = tile - blk * tpb, mi = tt / (b / 8), nj = tt - mi * (b / 8);
      const int s0 = blk * 2 * b;
      tile_product(X, s0, s0 + b, X, s0, s0, Y, s0, s0 + b, b, mi, nj, -1.0);           // X12 = -X11 Y12
    }
    __syncthreads();
  }
  for (int idx = tid; idx < n * n; idx += CI_THREADS) {
    const int i = idx / n, j = idx - i * n;
    R[(int64_t)i * ldr + j] = A[i * CI_P + j];
    Rinv[(int64_t)i * ldri + j] = X[i * CI_P + j];
  }
}


// G <- G - C^H C for the first pass of a 64-column block (Pythagorean step of BCGS-PIP): C is jl x b (ld ldc, the
// couplings of the block to the jl earlier columns), G is b x b Hermitian (ld ldg; only its UPPER triangle is
// read by the factorisation kernels, and only that is updated).  One CTA per upper 8 x 8 tile; its 16 warps split
// the jl rows of C (eight k-steps of loads in flight per warp), partial tiles are summed in warp order
// (deterministic): one launch instead of a split-K GEMM and its reduce kernel on the chain of every block.
constexpr int GC_WARPS = 16, GC_UNR = 8;
template <typename T>
__global__ void __launch_bounds__(GC_WARPS * 32) gram_correct_kernel(const T* C, int64_t ldc, int64_t jl, T* G, int64_t ldg, int b) {
  typedef Num<T> N_;
  constexpr bool CPLX = sizeof(T) == 16;
  griddep_wait();
  griddep_launch_dependents();
  const int nt = (b + 7) / 8;
  int mi = 0, rem = blockIdx.x;          // tile (mi <= nj) of the upper triangle, row-major
  while (rem >= nt - mi) { rem -= nt - mi; ++mi; }
  const int nj = mi + rem;
  const int I0 = mi * 8, K0 = nj * 8;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, gq = lane >> 2, tq = lane & 3;
  // rows of C per warp: a multiple of 4 (the DMMA k)
  const int64_t per = ((jl + GC_WARPS * 4 - 1) / (GC_WARPS * 4)) * 4;
  const int64_t p_beg = (int64_t)warp * per, p_end = (p_beg + per < jl) ? p_beg + per : jl;
  const bool ci_ok = I0 + gq < b, ck_ok = K0 + gq < b;
  double acc[CPLX ? 4 : 2] = {};
  for (int64_t p0 = p_beg; p0 < p_end; p0 += 4 * GC_UNR) {
    T a[GC_UNR], q[GC_UNR];
#pragma unroll
    for (int u = 0; u < GC_UNR; ++u) {
      const int64_t p = p0 + 4 * u + tq;
      a[u] = (p < p_end && ci_ok) ? N_::conj(C[p * ldc + I0 + gq]) : N_::zero();
      q[u] = (p < p_end && ck_ok) ? C[p * ldc + K0 + gq] : N_::zero();
    }
#pragma unroll
    for (int u = 0; u < GC_UNR; ++u) {
      if constexpr (CPLX) {
        dmma884(acc[0], acc[1], a[u].x, q[u].x);
        dmma884(acc[0], acc[1], -a[u].y, q[u].y);
        dmma884(acc[2], acc[3], a[u].x, q[u].y);
        dmma884(acc[2], acc[3], a[u].y, q[u].x);
      } else {
        dmma884(acc[0], acc[1], a[u], q[u]);
      }
    }
  }
  __shared__ T part[GC_WARPS][64];
  {
    T* d = &part[warp][gq * 8 + 2 * tq];
    if constexpr (CPLX) { d[0] = make_double2(acc[0], acc[2]); d[1] = make_double2(acc[1], acc[3]); }
    else { d[0] = acc[0]; d[1] = acc[1]; }
  }
  __syncthreads();
  if (tid < 64) {
    const int i = I0 + (tid >> 3), k = K0 + (tid & 7);
    if (i < b && k < b && k >= i) {
      T sum = part[0][tid];
#pragma unroll
      for (int w = 1; w < GC_WARPS; ++w) sum = N_::add(sum, part[w][tid]);
      G[(int64_t)i * ldg + k] = N_::sub(G[(int64_t)i * ldg + k], sum);
    }
  }
}

template <typename T>
__global__ void scale_by_kernel(T* x, int64_t n, const double* s) {
  const double f = s[0];
  const int64_t step = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += step) x[i] = Num<T>::scale(x[i], f);
}

// Side stream + events of qr_bcgs2 (the projection W - Q C of a first pass runs next to the 64 x 64 Cholesky
// kernel), one set per host thread and device (batch.py drives several streams from several host threads).
struct QrSide {
  int dev = -1;
  cudaStream_t s = nullptr;
  cudaEvent_t fork = nullptr, join = nullptr, fork2 = nullptr, join2 = nullptr;
  int get(int device) {
    if (dev == device && s) return 0;
    TNB_CUDA_CHECK(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking));
    TNB_CUDA_CHECK(cudaEventCreateWithFlags(&fork, cudaEventDisableTiming));
    TNB_CUDA_CHECK(cudaEventCreateWithFlags(&join, cudaEventDisableTiming));
    TNB_CUDA_CHECK(cudaEventCreateWithFlags(&fork2, cudaEventDisableTiming));
    TNB_CUDA_CHECK(cudaEventCreateWithFlags(&join2, cudaEventDisableTiming));
    dev = device;
    return 0;
  }
};
static thread_local QrSide g_qr_side;
constexpr int g_qr_overlap = TNB_EXP_QR_OVERLAP;

// G = I + E (n x n Hermitian, upper triangle read) with |E| <= tol entrywise  ->  R = I + U, Rinv = I - U,
// U = striu(E) + diag(E)/2: the Cholesky factor and its inverse to O(|E|^2).  Any entry outside tol (or a
// NaN) raises flag[0]; the outputs stay finite either way.  Plain grid-stride kernel, any n.
// C2 (jl x n, ld ldg, the rows above G: the couplings of the group to the earlier columns): when every entry of C2
// and of E is within skip_tol -- the group came out of its first pass orthonormal to rounding level -- skip[0]
// stays non-zero and the update kernels queued behind (SkipScope) return at once; any larger entry clears it.
// The skip predicate of a group's second pass, evaluated on what the S product left: C2 (jl x n, ld ldg, the couplings
// to the earlier columns) and G0 = W^H W (n x n, upper triangle).  Every entry of C2 and of G0 - I within skip_tol
// leaves skip[0] non-zero (set by the caller); any larger entry (or a NaN) clears it.  Runs BEFORE the Pythagorean
// correction G0 - C2^H C2, which is itself skipped with the rest (for a group that passes it is ~1e-28).
template <typename T>
__global__ void skip_predicate_kernel(const T* G0, int64_t ldg, int64_t n, const T* C2, int64_t jl, double skip_tol, int* skip) {
  typedef Num<T> N_;
  griddep_wait();
  griddep_launch_dependents();
  const int64_t step = (int64_t)gridDim.x * blockDim.x;
  const double t2 = skip_tol * skip_tol;
  bool keep = false;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < jl * n; idx += step) {
    const int64_t i = idx / n, j = idx - i * n;
    if (!(N_::abs2(C2[i * ldg + j]) <= t2)) keep = true;
  }
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < n * n; idx += step) {
    const int64_t i = idx / n, j = idx - i * n;
    if (j < i) continue;
    const T g = G0[i * ldg + j];
    const T e = (j == i) ? N_::sub(g, N_::one()) : g;
    if (!(N_::abs2(e) <= t2)) keep = true;
  }
  if (keep) skip[0] = 0;
}

template <typename T>
__global__ void near_identity_kernel(const T* G, int64_t ldg, int64_t n, T* R, int64_t ldr, T* Rinv, int64_t ldri,
                                     double tol, int* flag, const int* skip) {
  typedef Num<T> N_;
  griddep_wait();
  griddep_launch_dependents();
  if (skip && *skip) return;   // the group needs no second pass: nobody reads R / Rinv
  const int64_t step = (int64_t)gridDim.x * blockDim.x;
  bool bad = false;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < n * n; idx += step) {
    const int64_t i = idx / n, j = idx - i * n;
    T r = N_::zero(), ri = N_::zero();
    if (j > i) {
      const T e = G[i * ldg + j];
      if (!(N_::abs2(e) <= tol * tol)) bad = true;
      else { r = e; ri = N_::sub(N_::zero(), e); }
    } else if (j == i) {
      const T g = G[i * ldg + i];
      const T e = N_::sub(g, N_::one());
      double h = 0.0;
      if (!(N_::abs2(e) <= tol * tol)) bad = true;
      else h = 0.5 * (N_::real(g) - 1.0);
      r = N_::from(1.0 + h, 0.0);
      ri = N_::from(1.0 - h, 0.0);
    }
    R[i * ldr + j] = r;
    Rinv[i * ldri + j] = ri;
  }
  if (bad) flag[0] = 1;
}

// returns 0 on success, 1 when the device flag asks for the Householder path, < 0 on errors
template <typename T>
static int qr_bcgs2(int dtype, int64_t m, int64_t n, const void* A, int64_t lda, void* Q, void* R, void* ws,
                    int scale_mode, double** scale_out, bool lagged, cudaStream_t st) {
  typedef Num<T> N_;
  const QrLayout L = qr_layout(dtype, m, n);
  char* base = (char*)ws;
  const int64_t k = L.k;
  T* Wsc = (T*)(base + L.off_w);                      // scaled copy of the columns beyond k (wide matrices)
  T* Qb = Q ? (T*)Q : (T*)(base + L.off_v);           // m x k orthonormal factor
  const int64_t ldq = Q ? k : L.ldv;
  constexpr int64_t LDB = QR_GB;                      // pitch of P2, Sb, Bc and the small factors
  T* P2 = (T*)(base + L.off_p2);                      // m x QR_GB
  T* Sb = (T*)(base + L.off_cb);                      // k x QR_GB   [C2; G0] of the second pass
  T* Bc = (T*)(base + L.off_bc);                      // k x QR_GB   [-C R^-1; R^-1]
  T* R1 = (T*)(base + L.off_rt);                      // three QR_GB x QR_GB matrices
  T* R2 = R1 + QR_GB * QR_GB;
  T* Rt = R2 + QR_GB * QR_GB;
  T* Rg = (T*)R;
  void* sk = L.sk_bytes ? (void*)(base + L.off_sk) : nullptr;
  // first pass of the lagged scheme: W - Q C on a side stream next to the Cholesky kernel; the side GEMM gets
  // the last quarter of the split-K scratch, which also keeps its grid below one wave (an SM stays free for
  // the single-CTA Cholesky kernel)
  const bool overlap = lagged && g_qr_overlap && L.sk_bytes > 0;
  constexpr int side_pct = TNB_EXP_QR_SIDE_PCT;
  const size_t sk_main = overlap ? ((L.sk_bytes / 100) * (100 - side_pct)) & ~(size_t)255 : L.sk_bytes;
  void* sk_side = overlap ? (void*)(base + L.off_sk + sk_main) : nullptr;
  const size_t sk_side_bytes = overlap ? L.sk_bytes - sk_main : 0;
  QrSide& side = g_qr_side;
  if (overlap) {
    int device = 0;
    TNB_CUDA_CHECK(cudaGetDevice(&device));
    const int rs = side.get(device);
    if (rs) return rs;
  }
  int* flag = (int*)(base + L.off_sc + 128);
  int* skipf = flag + 1;   // second pass of a group: non-zero = the group needs no update (see near_identity_kernel)
  double* sc = nullptr;
  if (scale_mode != 0) {
    unsigned long long* bits = (unsigned long long*)(base + L.off_sc);
    sc = (double*)(base + L.off_sc) + 1;
    TNB_CUDA_CHECK(cudaMemsetAsync(bits, 0, sizeof(unsigned long long), st));
    amax_kernel<T><<<blocks_for(m * n), 256, 0, st>>>((const T*)A, lda, m, n, bits);
    TNB_LAUNCH_CHECK();
    pow2_scale_kernel<<<1, 1, 0, st>>>(bits, sc);
    TNB_LAUNCH_CHECK();
    if (scale_out) *scale_out = sc;
  }
  TNB_CUDA_CHECK(cudaMemsetAsync(flag, 0, sizeof(int), st));
  TNB_CUDA_CHECK(cudaMemsetAsync(Rg, 0, (size_t)k * n * sizeof(T), st));
  auto kern = chol_inv_kernel<T>;
  constexpr size_t ci_smem = (size_t)3 * CI_N * CI_P * sizeof(T);
  static PerDeviceOnce once;
  if (once.need()) {
    TNB_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ci_smem));
    once.done();
  }
  // the small factor of a pass: G (ld ldg) -> Rout (ld ldr), R^-1 -> Rinv (ld LDB)
  struct Factor { int kind; T* Rout; int64_t ldr; double rel_floor, abs_floor, near_tol; };
  auto factorise = [&](const Factor& f, const T* G, int64_t ldg, T* Rinv, int64_t b, int64_t jl) -> int {
    ProfScope prof(KC_QR_PANEL, st, (sizeof(T) == 16 ? 4.0 : 1.0) * 2.0 / 3.0 * (double)b * b * b);
    if (f.kind == 0) {
      TNB_CUDA_CHECK(launch_k(kern, dim3(1), dim3(CI_THREADS), ci_smem, st, G, ldg, (int)b, f.Rout, f.ldr, Rinv, LDB, f.rel_floor,
                              f.abs_floor, f.near_tol, flag));
    } else {
      TNB_CUDA_CHECK(launch_k(near_identity_kernel<T>, dim3(blocks_for(b * b)), dim3(256), 0, st, G, ldg, b, f.Rout,
                              f.ldr, Rinv, LDB, f.near_tol, flag, (const int*)skipf));
    }
    TNB_LAUNCH_CHECK();
    count_launch();
    return 0;
  };
  int rc;
  // one pass over the panel Qp = Q[:, j0 : j0 + b] against the columns Qq = Q[:, q0 : j0] (jl = j0 - q0 of them):
  //   S (ld lds) <- [Qq, W]^H W;  G = G0 - C^H C in place;  factor;  Bc = [-C R^-1; R^-1];  W <- [Qq, W] Bc
  auto pass = [&](int64_t q0, int64_t j0, int64_t b, T* Qp, T* S, int64_t lds, const Factor& f) -> int {
    const int64_t jl = j0 - q0, kk = jl + b;
    const T* Qq = Qb + q0;
    int r_ = gemm_ws(dtype, TNB_OP_C, TNB_OP_N, kk, b, m, 1, 0, Qq, ldq, 0, Qp, ldq, 0, 0, 0, S, lds, 0, 1, sk, sk_main, st);
    if (r_) return r_;
    T* G = S + jl * lds;
    T* Ri = Bc + jl * LDB;
    const bool fork = overlap && f.kind == 0 && jl > 0;
    if (fork) {
      // side stream: W <- W - Qq C in place (C is complete; the Cholesky does not need W)
      TNB_CUDA_CHECK(cudaEventRecord(side.fork, st));
      TNB_CUDA_CHECK(cudaStreamWaitEvent(side.s, side.fork, 0));
    }
    if (f.kind == 1) {
      // second pass of a group: decide on the device whether there is anything to correct; the correction of G, the
      // factor and every update kernel behind them return at once when there is not
      TNB_CUDA_CHECK(launch_k(skip_predicate_kernel<T>, dim3(blocks_for((jl + b) * b)), dim3(256), 0, st, (const T*)G, lds, b,
                              (const T*)S, jl, QR_SKIP_TOL, skipf));
      TNB_LAUNCH_CHECK();
    }
    bool corrected = (jl == 0);
#ifndef TNB_EXP_QR_GEMM_GCORR
    if (!corrected && f.kind == 0 && b <= QR_CB) {
      // first pass of a block: the 64 x 64 correction in one launch (only the upper triangle, which is all chol_inv_kernel reads)
      ProfScope prof(KC_GEMM, st, (sizeof(T) == 16 ? 8.0 : 2.0) * (double)b * (double)b * (double)jl * 0.5);
      const int nt = (int)((b + 7) / 8);
      TNB_CUDA_CHECK(launch_k(gram_correct_kernel<T>, dim3((unsigned)(nt * (nt + 1) / 2)), dim3(GC_WARPS * 32), 0, st, (const T*)S,
                              lds, jl, G, lds, (int)b));
      TNB_LAUNCH_CHECK();
      corrected = true;
    }
#endif
    if (!corrected) {
      SkipScope skip_scope(f.kind == 1 ? skipf : nullptr);
      r_ = gemm_ws(dtype, TNB_OP_C, TNB_OP_N, b, b, jl, -1, 0, S, lds, 0, S, lds, 0, 1, 0, G, lds, 0, 1, sk, sk_main, st);
      if (r_) return r_;
    }
    factorise(f, G, lds, Ri, b, jl);
    if (fork) {
      r_ = gemm_ws(dtype, TNB_OP_N, TNB_OP_N, m, b, jl, -1, 0, Qq, ldq, 0, S, lds, 0, 1, 0, Qp, ldq, 0, 1, sk_side, sk_side_bytes,
                   side.s);
      TNB_CUDA_CHECK(cudaEventRecord(side.join, side.s));
      TNB_CUDA_CHECK(cudaStreamWaitEvent(st, side.join, 0));
      if (r_) return r_;
      // W <- (W - Qq C) R^-1, in place
      if (b <= PS_N) {
        TNB_CUDA_CHECK(launch_k(panel_scale_kernel<T>, dim3((unsigned)((m + PS_ROWS - 1) / PS_ROWS)), dim3(256), 0, st,
                                Qp, ldq, m, (int)b, (const T*)Ri, LDB));
        TNB_LAUNCH_CHECK();
        return 0;
      }
      r_ = gemm(dtype, TNB_OP_N, TNB_OP_N, m, b, b, 1, 0, Qp, ldq, 0, Ri, LDB, 0, 0, 0, P2, LDB, 0, 1, st);
      if (r_) return r_;
    } else {
      // (second pass of a group: these kernels return at once when the device predicate found nothing to correct)
      SkipScope skip_scope(f.kind == 1 ? skipf : nullptr);
      if (jl > 0) {
        r_ = gemm(dtype, TNB_OP_N, TNB_OP_N, jl, b, b, -1, 0, S, lds, 0, Ri, LDB, 0, 0, 0, Bc, LDB, 0, 1, st);
        if (r_) return r_;
      }
      r_ = gemm_ws(dtype, TNB_OP_N, TNB_OP_N, m, b, kk, 1, 0, Qq, ldq, 0, Bc, LDB, 0, 0, 0, P2, LDB, 0, 1, sk, sk_main, st);
      if (r_) return r_;
      TNB_CUDA_CHECK(launch_k(copy2d_kernel<T>, dim3(blocks_for(m * b)), dim3(256), 0, st, (const T*)P2, LDB, Qp, ldq, m, b,
                              (const double*)nullptr, (const int*)g_skip_flag));
      TNB_LAUNCH_CHECK();
      return 0;
    }
    TNB_CUDA_CHECK(launch_k(copy2d_kernel<T>, dim3(blocks_for(m * b)), dim3(256), 0, st, (const T*)P2, LDB, Qp, ldq, m, b,
                            (const double*)nullptr, (const int*)nullptr));
    TNB_LAUNCH_CHECK();
    return 0;
  };
  if (!lagged) {
    // two passes per 64-column block
    for (int64_t j0 = 0; j0 < k; j0 += QR_CB) {
      const int64_t bj = (k - j0 < QR_CB) ? (k - j0) : QR_CB;
      T* Qp = Qb + j0;
      T* Rj = Rg + j0;                // R[0:j0, J], ld n  (pass 1 leaves C1 there, and G0 -> G in R[J, J])
      T* Rjj = Rg + j0 * n + j0;
      copy2d_kernel<T><<<blocks_for(m * bj), 256, 0, st>>>((const T*)A + j0, lda, Qp, ldq, m, bj, sc);
      TNB_LAUNCH_CHECK();
      rc = pass(0, j0, bj, Qp, Rj, n, Factor{0, R1, LDB, 1e-10, 0.0, 0.0});
      if (rc) return rc;
      rc = pass(0, j0, bj, Qp, Sb, LDB, Factor{0, R2, LDB, 0.0, 0.25, 1e-8});
      if (rc) return rc;
      if (j0 > 0) {  // R[:j0, J] = C1 + C2 R1
        rc = gemm(dtype, TNB_OP_N, TNB_OP_N, j0, bj, bj, 1, 0, Sb, LDB, 0, R1, LDB, 0, 1, 0, Rj, n, 0, 1, st);
        if (rc) return rc;
      }
      rc = gemm(dtype, TNB_OP_N, TNB_OP_N, bj, bj, bj, 1, 0, R2, LDB, 0, R1, LDB, 0, 0, 0, Rjj, n, 0, 1, st);
      if (rc) return rc;
    }
  } else {
    // Lagged reorthogonalisation: the first pass runs block by block (64 columns, Cholesky in the small
    // kernel, R1 and the couplings C1 land directly in R); the second pass runs once per GROUP of 256
    // columns.  After one pass the group is orthonormal to ~eps cond^2, so its Gram matrix against
    // [earlier groups, itself] is I + E with tiny E and the first-order factor R2 = I + U serves for any
    // width -- a quarter of the second-pass launches, and GEMMs with N = 256.  |E| > 1e-8 anywhere raises the
    // flag (the caller then repeats the factorisation with two passes per block).
    bool rfix_pending = false;
    for (int64_t g0 = 0; g0 < k; g0 += QR_GB) {
      const int64_t bg = (k - g0 < QR_GB) ? (k - g0) : QR_GB;
      copy2d_kernel<T><<<blocks_for(m * bg), 256, 0, st>>>((const T*)A + g0, lda, Qb + g0, ldq, m, bg, sc);
      TNB_LAUNCH_CHECK();
      const int64_t q0 = 0;   // (a group-level projection first, q0 = g0, was measured slower: the Cholesky then sits on the chain)
      for (int64_t j0 = g0; j0 < g0 + bg; j0 += QR_CB) {
        const int64_t bj = (g0 + bg - j0 < QR_CB) ? (g0 + bg - j0) : QR_CB;
        rc = pass(q0, j0, bj, Qb + j0, Rg + q0 * n + j0, n, Factor{0, Rg + j0 * n + j0, n, 1e-10, 0.0, 0.0});
        if (rc) return rc;
      }
      if (rfix_pending) {   // the previous group's R update still reads Sb / R2 / Rt, which this pass overwrites
        TNB_CUDA_CHECK(cudaStreamWaitEvent(st, side.join2, 0));
        rfix_pending = false;
      }
      TNB_CUDA_CHECK(cudaMemsetAsync(skipf, 1, sizeof(int), st));   // non-zero until an entry beyond QR_SKIP_TOL clears it
      rc = pass(0, g0, bg, Qb + g0, Sb, LDB, Factor{1, R2, LDB, 0.0, 0.0, 1e-8});
      if (rc) return rc;
      // R[G, G] = R2 R1g,  R[:g0, G] = C1 + C2 R1g   (R1g = the group's first-pass factor, now in R[G, G]).
      // Nothing of the factorisation reads these entries of R again, so the three short kernels run on the side
      // stream, next to the first blocks of the following group (they cost ~60 us per group on the main stream).
      cudaStream_t sr = st;
      if (overlap) {
        TNB_CUDA_CHECK(cudaEventRecord(side.fork2, st));
        TNB_CUDA_CHECK(cudaStreamWaitEvent(side.s, side.fork2, 0));
        sr = side.s;
      }
      T* Rgg = Rg + g0 * n + g0;
      {
        SkipScope skip_scope(skipf);   // an untouched group keeps its first-pass R
        copy2d_kernel<T><<<blocks_for(bg * bg), 256, 0, sr>>>(Rgg, n, Rt, LDB, bg, bg, nullptr, skipf);
        TNB_LAUNCH_CHECK();
        rc = gemm(dtype, TNB_OP_N, TNB_OP_N, bg, bg, bg, 1, 0, R2, LDB, 0, Rt, LDB, 0, 0, 0, Rgg, n, 0, 1, sr);
        if (rc) return rc;
        if (g0 > 0) {
          rc = gemm(dtype, TNB_OP_N, TNB_OP_N, g0, bg, bg, 1, 0, Sb, LDB, 0, Rt, LDB, 0, 1, 0, Rg + g0, n, 0, 1, sr);
          if (rc) return rc;
        }
      }
      if (overlap) {
        TNB_CUDA_CHECK(cudaEventRecord(side.join2, side.s));
        rfix_pending = true;
      }
    }
    if (rfix_pending) TNB_CUDA_CHECK(cudaStreamWaitEvent(st, side.join2, 0));
  }
  if (n > k) {  // wide: R[:, k:] = Q^H (scale * A[:, k:])
    copy2d_kernel<T><<<blocks_for(m * (n - k)), 256, 0, st>>>((const T*)A + k, lda, Wsc, L.ldw, m, n - k, sc);
    TNB_LAUNCH_CHECK();
    rc = gemm_ws(dtype, TNB_OP_C, TNB_OP_N, k, n - k, m, 1, 0, Qb, ldq, 0, Wsc, L.ldw, 0, 0, 0, Rg + k, n, 0, 1, sk, sk_main, st);
    if (rc) return rc;
  }
  if (scale_mode == 1) {
    scale_by_kernel<T><<<blocks_for(k * n), 256, 0, st>>>(Rg, k * n, sc + 1);
    TNB_LAUNCH_CHECK();
  }
  int h = 0;
  TNB_CUDA_CHECK(cudaMemcpyAsync(&h, flag, sizeof(int), cudaMemcpyDeviceToHost, st));
  TNB_CUDA_CHECK(cudaStreamSynchronize(st));
  (void)sizeof(N_);
  return h ? 1 : 0;
}

// scale_mode 0: factor A as it is; 1: factor 2^e A (e from max|A|, so no dot product can overflow) and
// return R of A itself; 2: as 1 but return R of the SCALED matrix and the factors in scale_out[0..1]
// (scale, 1/scale) -- used by svd.cu, which folds 1/scale into the singular values.
//
// Two-level blocking: columns are factored in inner panels of QR_NB = 32 (cluster kernel), grouped in
// outer blocks of QR_NBO = 128.  Inside an outer block each panel's reflector is applied to the rest
// of the block only; the block's 128 x 128 T factor is assembled from the panels' T factors
// (T12 = -T11 (V1^H V2) T22), and the trailing matrix / the explicit Q see one K = 128 compact-WY
// update per outer block.  The long-K products V^H A use split-K.
template <typename T>
static int qr_impl(int dtype, int64_t m, int64_t n, const void* A, int64_t lda, void* Q, void* R, void* ws,
                   int scale_mode, double** scale_out, cudaStream_t st) {
  const QrLayout L = qr_layout(dtype, m, n);
  char* base = (char*)ws;
  T* W = (T*)(base + L.off_w);
  T* V = (T*)(base + L.off_v);
  T* Tb = (T*)(base + L.off_t);
  T* W1 = (T*)(base + L.off_w1);
  T* W2 = (T*)(base + L.off_w2);
  void* sk = L.sk_bytes ? (void*)(base + L.off_sk) : nullptr;
  const int64_t k = L.k;
  double* sc = nullptr;
  if (scale_mode != 0) {
    unsigned long long* bits = (unsigned long long*)(base + L.off_sc);
    sc = (double*)(base + L.off_sc) + 1;
    TNB_CUDA_CHECK(cudaMemsetAsync(bits, 0, sizeof(unsigned long long), st));
    amax_kernel<T><<<blocks_for(m * n), 256, 0, st>>>((const T*)A, lda, m, n, bits);
    TNB_LAUNCH_CHECK();
    pow2_scale_kernel<<<1, 1, 0, st>>>(bits, sc);
    TNB_LAUNCH_CHECK();
    if (scale_out) *scale_out = sc;
  }
  copy2d_kernel<T><<<blocks_for(m * n), 256, 0, st>>>((const T*)A, lda, W, L.ldw, m, n, sc);
  TNB_LAUNCH_CHECK();
  // V must be zero outside the panels' own columns/rows it writes (rows above a panel); T blocks start zero
  TNB_CUDA_CHECK(cudaMemsetAsync(V, 0, (size_t)m * L.ldv * sizeof(T), st));
  TNB_CUDA_CHECK(cudaMemsetAsync(Tb, 0, (size_t)L.nouter * QR_NBO * QR_NBO * sizeof(T), st));
  int rc;
  for (int64_t ob = 0; ob < L.nouter; ++ob) {
    const int64_t j0 = ob * QR_NBO;
    const int64_t jbo = (k - j0 < QR_NBO) ? (k - j0) : QR_NBO;
    T* To = Tb + ob * QR_NBO * QR_NBO;  // ld = QR_NBO
    const T* Vo = V + j0 * L.ldv + j0;
    const int64_t mro = m - j0;
    for (int64_t cb = 0; cb < jbo; cb += QR_NB) {
      const int64_t i0 = j0 + cb;
      const int jb = (int)((jbo - cb < QR_NB) ? (jbo - cb) : QR_NB);
      QrPanelArgs a;
      a.W = W; a.V = V; a.T = To + cb * QR_NBO + cb; a.ldt = QR_NBO;
      a.m = m; a.ldw = L.ldw; a.ldv = L.ldv; a.j0 = i0; a.jb = jb;
      a.fast = g_qr_fast_panel;
      rc = launch_panel<T>(a, st);
      if (rc) return rc;
      const T* Vp = V + i0 * L.ldv + i0;
      const T* Tp = To + cb * QR_NBO + cb;
      const int64_t mri = m - i0;
      const int64_t nin = jbo - cb - jb;  // columns of this outer block still to the right
      if (nin > 0) {
        T* Ain = W + i0 * L.ldw + i0 + jb;
        // Ain -= Vp Tp^H Vp^H Ain
        rc = gemm_ws(dtype, TNB_OP_C, TNB_OP_N, jb, nin, mri, 1, 0, Vp, L.ldv, 0, Ain, L.ldw, 0, 0, 0, W1, nin, 0, 1, sk, L.sk_bytes, st);
        if (rc) return rc;
        rc = gemm(dtype, TNB_OP_C, TNB_OP_N, jb, nin, jb, 1, 0, Tp, QR_NBO, 0, W1, nin, 0, 0, 0, W2, nin, 0, 1, st);
        if (rc) return rc;
        rc = gemm(dtype, TNB_OP_N, TNB_OP_N, mri, nin, jb, -1, 0, Vp, L.ldv, 0, W2, nin, 0, 1, 0, Ain, L.ldw, 0, 1, st);
        if (rc) return rc;
      }
      if (cb > 0) {
        // To[0:cb, cb:cb+jb] = -To[0:cb,0:cb] (V1^H V2) Tp,  V1 = Vo[:, 0:cb], V2 = Vo[:, cb:cb+jb]
        rc = gemm_ws(dtype, TNB_OP_C, TNB_OP_N, cb, jb, mro, 1, 0, Vo, L.ldv, 0, Vo + cb, L.ldv, 0, 0, 0, W1, jb, 0, 1, sk, L.sk_bytes, st);
        if (rc) return rc;
        rc = gemm(dtype, TNB_OP_N, TNB_OP_N, cb, jb, jb, 1, 0, W1, jb, 0, Tp, QR_NBO, 0, 0, 0, W2, jb, 0, 1, st);
        if (rc) return rc;
        rc = gemm(dtype, TNB_OP_N, TNB_OP_N, cb, jb, cb, -1, 0, To, QR_NBO, 0, W2, jb, 0, 0, 0, To + cb, QR_NBO, 0, 1, st);
        if (rc) return rc;
      }
    }
    const int64_t n2 = n - j0 - jbo;
    if (n2 > 0) {
      T* A2 = W + j0 * L.ldw + j0 + jbo;
      // W1 = Vo^H A2 ; W2 = To^H W1 ; A2 -= Vo W2
      rc = gemm_ws(dtype, TNB_OP_C, TNB_OP_N, jbo, n2, mro, 1, 0, Vo, L.ldv, 0, A2, L.ldw, 0, 0, 0, W1, n2, 0, 1, sk, L.sk_bytes, st);
      if (rc) return rc;
      rc = gemm(dtype, TNB_OP_C, TNB_OP_N, jbo, n2, jbo, 1, 0, To, QR_NBO, 0, W1, n2, 0, 0, 0, W2, n2, 0, 1, st);
      if (rc) return rc;
      rc = gemm(dtype, TNB_OP_N, TNB_OP_N, mro, n2, jbo, -1, 0, Vo, L.ldv, 0, W2, n2, 0, 1, 0, A2, L.ldw, 0, 1, st);
      if (rc) return rc;
    }
  }
  if (R) {
    extract_r_kernel<T><<<blocks_for(k * n), 256, 0, st>>>(W, L.ldw, (T*)R, k, n, scale_mode == 1 ? sc : nullptr);
    TNB_LAUNCH_CHECK();
  }
  if (Q) {
    T* Qo = (T*)Q;
    eye_kernel<T><<<blocks_for(m * k), 256, 0, st>>>(Qo, m, k, k);
    TNB_LAUNCH_CHECK();
    for (int64_t ob = L.nouter - 1; ob >= 0; --ob) {
      const int64_t j0 = ob * QR_NBO;
      const int64_t jbo = (k - j0 < QR_NBO) ? (k - j0) : QR_NBO;
      const int64_t nc = k - j0, mro = m - j0;
      const T* Vo = V + j0 * L.ldv + j0;
      const T* To = Tb + ob * QR_NBO * QR_NBO;
      T* Qs = Qo + j0 * k + j0;
      // Qs := (I - Vo To Vo^H) Qs
      rc = gemm_ws(dtype, TNB_OP_C, TNB_OP_N, jbo, nc, mro, 1, 0, Vo, L.ldv, 0, Qs, k, 0, 0, 0, W1, nc, 0, 1, sk, L.sk_bytes, st);
      if (rc) return rc;
      rc = gemm(dtype, TNB_OP_N, TNB_OP_N, jbo, nc, jbo, 1, 0, To, QR_NBO, 0, W1, nc, 0, 0, 0, W2, nc, 0, 1, st);
      if (rc) return rc;
      rc = gemm(dtype, TNB_OP_N, TNB_OP_N, mro, nc, jbo, -1, 0, Vo, L.ldv, 0, W2, nc, 0, 1, 0, Qs, k, 0, 1, st);
      if (rc) return rc;
    }
  }
  return 0;
}

// internal entry used by svd.cu as well
constexpr int g_qr_bcgs = TNB_EXP_QR_BCGS;

int qr(int dtype, int64_t m, int64_t n, const void* A, int64_t lda, void* Q, void* R, void* ws, int scale_mode,
       double** scale_out, cudaStream_t st) {
  // many columns: Gram-Schmidt path (all GEMMs); falls back to Householder when its device checks trip
  const int64_t k = m < n ? m : n;
  if (g_qr_bcgs && R && k >= 2 * QR_CB) {
    int rc = 1;
    // first the lagged second pass, then two passes per block, then Householder
    for (int lag = (g_qr_bcgs >= 2 && k > 2 * QR_CB) ? 1 : 0; lag >= 0 && rc == 1; --lag) {
      rc = (dtype == TNB_F64) ? qr_bcgs2<double>(dtype, m, n, A, lda, Q, R, ws, scale_mode, scale_out, lag != 0, st)
                              : qr_bcgs2<cplx>(dtype, m, n, A, lda, Q, R, ws, scale_mode, scale_out, lag != 0, st);
    }
    if (rc <= 0) return rc;
  }
  if (dtype == TNB_F64) return qr_impl<double>(dtype, m, n, A, lda, Q, R, ws, scale_mode, scale_out, st);
  return qr_impl<cplx>(dtype, m, n, A, lda, Q, R, ws, scale_mode, scale_out, st);
}
size_t qr_workspace(int dtype, int64_t m, int64_t n) { return qr_layout(dtype, m, n).total; }

}  // namespace tnb

extern "C" size_t tnb_qr_workspace(int dtype, int64_t m, int64_t n) {
  if (m <= 0 || n <= 0 || (dtype != TNB_F64 && dtype != TNB_C128)) return 0;
  return tnb::qr_workspace(dtype, m, n);
}

extern "C" int tnb_qr(int dtype, int64_t m, int64_t n, const void* A, int64_t lda, void* Q, void* R, void* ws,
                      size_t ws_bytes, void* stream) {
  if (dtype != TNB_F64 && dtype != TNB_C128) return TNB_E_ARG;
  if (m < 0 || n < 0 || lda < n) return TNB_E_ARG;
  if (m == 0 || n == 0) return 0;
  if (!A || !ws) return TNB_E_ARG;
  if (ws_bytes < tnb::qr_workspace(dtype, m, n)) return TNB_E_WORKSPACE;
  return tnb::qr(dtype, m, n, A, lda, Q, R, ws, 1, nullptr, (cudaStream_t)stream);
}

#ifdef TNB_EXP_STAMPS
// kernel experiments: phase timestamps (clock64) of the last fast-path panel, see QR_STAMP
extern "C" int tnb_debug_qr_stamps(long long* out, int n) {
  if (!out || n <= 0 || n > 32) return TNB_E_ARG;
  TNB_CUDA_CHECK(cudaMemcpyFromSymbol(out, tnb::g_qr_dbg, (size_t)n * sizeof(long long)));
  return 0;
}
#endif
