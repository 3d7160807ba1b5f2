// FP64 / complex128 GEMM on the sm_100a FP64 tensor pipe (DMMA.8x8x4).
//
//   C[b] = alpha * op(A[b]) * op(B[b]) + beta * C[b]        (row-major)
//
// tcgen05.mma has no f64 kind, so FP64 tensor math on Blackwell is
// mma.sync.m8n8k4.f64 (one DMMA.8x8x4 per instruction).  Layout of one CTA:
//   * real:    128x128 C tile, 8 warps (2 x 4), warp tile 64x32, BK = 16
//   * complex: 128x64  C tile, 8 warps (4 x 2), warp tile 32x32, BK = 8,
//              3 DMMAs per (m8,n8,k4) tile pair (3-multiplication form):
//              P1 += ar*br, P2 += ai*bi, P3 += (ar+ai)*(br+bi); re = P1 - P2,
//              im = P3 - P1 - P2 in the epilogue (interleaved re/im stays
//              interleaved in shared memory; fragments are 16-byte LDS)
//   * a "small" variant (64x64 real / 64x32 complex, 4 warps) for skinny shapes
// Operands are staged by a 4-stage cp.async (LDGSTS, 16-byte, zero-fill at the
// edges) pipeline.  Either operand may be K-contiguous ([mn][k]) or
// MN-contiguous ([k][mn]) in global memory -- this is how the label->axis
// permutation of np.tensordot is folded into operand staging instead of being
// materialised -- and the shared-memory row pitch is padded so that the
// per-lane fragment loads (row = lane/4, k = lane%4) are bank-conflict free:
//   K-contiguous : pitch BK+4 (real) / BK+4 (complex)
//   MN-contiguous: pitch BMN+4 (real) / BMN+2 (complex)
// Conjugation of either operand is a sign flip on the imaginary fragment.
//
// The full-size tiles (the ones that carry the sweeps) are fed by TMA instead (gemm_tma_kernel for
// complex128, gemm_tma_real_kernel for float64: 28.3 -> 33.1 TFLOP/s on 8192^3): cp.async.bulk.tensor copies through two tensor maps (SASS UTMALDG), issued by
// one thread, completing on per-stage mbarriers; six stages of dense 128-byte-swizzled tiles (no
// padding: 24 KB per stage instead of 36).  Bank conflicts are avoided by the hardware swizzle plus
// a fixed permutation of the rows inside every group of 8 (see sigma8), edges by the zero fill of
// out-of-bounds boxes.  Anything a tensor map cannot describe (zero or unaligned strides, odd
// leading dimensions of float64 operands) and the small tiles take the cp.async kernel.
// Algorithmic work: 2*M*N*K flop (real), 8*M*N*K flop (complex).
#include <cuda.h>

#include "common.cuh"

namespace tnb {

struct GemmArgs {
  const void* A;
  const void* B;
  void* C;
  int64_t M, N, K;
  int64_t lda, ldb, ldc;
  int64_t sA, sB, sC;  // batch strides (elements)
  double alpha_r, alpha_i, beta_r, beta_i;
  int conjA, conjB;
  int tiles_m, tiles_n;
  // split-K: blockIdx.z owns K range [z*k_per_split, ...) and writes its raw partial product
  // (alpha = 1, beta = 0) to part + z*M*N (row-major, ld = N); splitk_reduce_kernel finishes.
  int splits;
  int64_t k_per_split;
  void* part;
  const int* skip;   // device flag: the kernel returns at once when *skip != 0 (SkipScope), or null
};

template <bool CPLX, bool SMALL> struct Cfg;
template <> struct Cfg<false, false> { static constexpr int BM = 128, BN = 128, BK = 16, WM = 64, WN = 32, WARPS_M = 2, WARPS_N = 4; };
template <> struct Cfg<false, true>  { static constexpr int BM = 64,  BN = 64,  BK = 16, WM = 32, WN = 32, WARPS_M = 2, WARPS_N = 2; };
template <> struct Cfg<true, false>  { static constexpr int BM = 128, BN = 64,  BK = 8,  WM = 32, WN = 32, WARPS_M = 4, WARPS_N = 2; };
template <> struct Cfg<true, true>   { static constexpr int BM = 64,  BN = 32,  BK = 8,  WM = 32, WN = 16, WARPS_M = 2, WARPS_N = 2; };

constexpr int STAGES = 4;

template <bool CPLX, bool KC, int BMN, int BK> struct TileLayout {
  // pitch in elements
  static constexpr int PITCH = KC ? (BK + 4) : (BMN + (CPLX ? 2 : 4));
  static constexpr int ROWS = KC ? BMN : BK;
  static constexpr int ELEMS = ROWS * PITCH;
};

// Stage one operand tile (BMN x BK logical) into shared memory.
//   KC  : global is [mn][k] with k contiguous  -> smem [mn][k]
//   !KC : global is [k][mn] with mn contiguous -> smem [k][mn]
// VEC (real only): 16-byte packets of two doubles; requires even ld / offsets
// and a 16-byte aligned base.  Complex elements are always 16-byte packets.
template <typename T, bool CPLX, bool KC, bool VEC, int BMN, int BK, int THREADS>
__device__ __forceinline__ void load_tile(T* smem, const T* __restrict__ g, int64_t ld, int64_t mn0, int64_t k0,
                                          int64_t MN, int64_t K, int tid) {
  typedef TileLayout<CPLX, KC, BMN, BK> L;
  constexpr int EPP = CPLX ? 1 : (VEC ? 2 : 1);  // elements per packet
  constexpr int INNER = KC ? BK : BMN;           // contiguous extent
  constexpr int OUTER = KC ? BMN : BK;
  constexpr int PPR = INNER / EPP;               // packets per row
  constexpr int TOTAL = OUTER * PPR;
  const int64_t in_lim = KC ? K : MN, out_lim = KC ? MN : K;
  const int64_t in0 = KC ? k0 : mn0, out0 = KC ? mn0 : k0;
#pragma unroll
  for (int it = 0; it < (TOTAL + THREADS - 1) / THREADS; ++it) {
    const int c = tid + it * THREADS;
    if (TOTAL % THREADS != 0 && c >= TOTAL) break;
    const int r = c / PPR, p = (c - r * PPR) * EPP;
    const int64_t go = out0 + r, gi = in0 + p;
    T* dst = smem + r * L::PITCH + p;
    int64_t valid = (go < out_lim) ? (in_lim - gi) : 0;
    if (valid > EPP) valid = EPP;
    if (valid < 0) valid = 0;
    // keep the source address in-bounds even when nothing is copied
    const T* src = (valid > 0) ? (g + go * ld + gi) : g;
    if constexpr (sizeof(T) * EPP == 16) cp_async16(dst, src, (int)valid * (int)sizeof(T));
    else cp_async8(dst, src, (int)valid * (int)sizeof(T));
  }
}

template <bool CPLX, bool SMALL, bool A_KC, bool B_KC, bool VEC>
__global__ void __launch_bounds__(Cfg<CPLX, SMALL>::WARPS_M * Cfg<CPLX, SMALL>::WARPS_N * 32)
gemm_kernel(GemmArgs g) {
  typedef Cfg<CPLX, SMALL> C_;
  typedef typename std::conditional<CPLX, double2, double>::type T;
  constexpr int BM = C_::BM, BN = C_::BN, BK = C_::BK, WM = C_::WM, WN = C_::WN;
  constexpr int THREADS = C_::WARPS_M * C_::WARPS_N * 32;
  constexpr int MT = WM / 8, NT = WN / 8;
  typedef TileLayout<CPLX, A_KC, BM, BK> LA;
  typedef TileLayout<CPLX, B_KC, BN, BK> LB;

  extern __shared__ __align__(16) unsigned char smem_raw[];
  T* sA = reinterpret_cast<T*>(smem_raw);
  T* sB = sA + STAGES * LA::ELEMS;
  griddep_wait();               // no-ops unless launched with programmatic stream serialization (launch_k)
  griddep_launch_dependents();
  if (g.skip && *g.skip) return;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int gq = lane >> 2, tq = lane & 3;  // fragment row / k index
  const int wm = warp / C_::WARPS_N, wn = warp % C_::WARPS_N;

  // tile coordinates: consecutive CTAs walk down M first so that the B tile
  // (shared by a column of C tiles) stays hot in L2
  const int tile = blockIdx.x;
  const int tm = tile % g.tiles_m, tn = tile / g.tiles_m;
  const int64_t m0 = (int64_t)tm * BM, n0 = (int64_t)tn * BN;
  const int64_t bz = blockIdx.y;
  const T* A = reinterpret_cast<const T*>(g.A) + bz * g.sA;
  const T* B = reinterpret_cast<const T*>(g.B) + bz * g.sB;
  T* Cg = reinterpret_cast<T*>(g.C) + bz * g.sC;
  const int64_t kbeg = (int64_t)blockIdx.z * g.k_per_split;
  const int64_t kend = (g.splits > 1 && kbeg + g.k_per_split < g.K) ? kbeg + g.k_per_split : g.K;
  int64_t ldc = g.ldc;
  if (g.splits > 1) {
    Cg = reinterpret_cast<T*>(g.part) + (int64_t)blockIdx.z * g.M * g.N;
    ldc = g.N;
  }

  // complex: 3-multiplication form, three accumulator pairs per 8x8 tile.  With a = ar + i sA ai, b = br + i sB bi
  // (sA, sB = -1 for a conjugated operand):
  //   P1 = ar br, P2 = ai bi, P3 = (ar + sA ai)(br + sB bi);  re = P1 - sA sB P2, im = P3 - P1 - sA sB P2
  double acc[MT][NT][CPLX ? 6 : 2];
#pragma unroll
  for (int i = 0; i < MT; ++i)
#pragma unroll
    for (int j = 0; j < NT; ++j)
#pragma unroll
      for (int r = 0; r < (CPLX ? 6 : 2); ++r) acc[i][j][r] = 0.0;

  const int KT = (int)((kend - kbeg + BK - 1) / BK);

  auto issue = [&](int kt) {
    if (kt < KT) {
      const int s = kt % STAGES;
      load_tile<T, CPLX, A_KC, VEC, BM, BK, THREADS>(sA + s * LA::ELEMS, A, g.lda, m0, kbeg + (int64_t)kt * BK, g.M, kend, tid);
      load_tile<T, CPLX, B_KC, VEC, BN, BK, THREADS>(sB + s * LB::ELEMS, B, g.ldb, n0, kbeg + (int64_t)kt * BK, g.N, kend, tid);
    }
    cp_async_commit();
  };

#pragma unroll
  for (int s = 0; s < STAGES - 1; ++s) issue(s);

  const double sgnA = g.conjA ? -1.0 : 1.0, sgnB = g.conjB ? -1.0 : 1.0;

  for (int kt = 0; kt < KT; ++kt) {
    cp_async_wait<STAGES - 2>();
    __syncthreads();
    issue(kt + STAGES - 1);
    const T* a_s = sA + (kt % STAGES) * LA::ELEMS;
    const T* b_s = sB + (kt % STAGES) * LB::ELEMS;
#pragma unroll
    for (int kk = 0; kk < BK; kk += 4) {
      T af[MT], bf[NT];
#pragma unroll
      for (int i = 0; i < MT; ++i) {
        const int m = wm * WM + i * 8 + gq, k = kk + tq;
        af[i] = A_KC ? a_s[m * LA::PITCH + k] : a_s[k * LA::PITCH + m];
      }
#pragma unroll
      for (int j = 0; j < NT; ++j) {
        const int n = wn * WN + j * 8 + gq, k = kk + tq;
        bf[j] = B_KC ? b_s[n * LB::PITCH + k] : b_s[k * LB::PITCH + n];
      }
      if constexpr (!CPLX) {
#pragma unroll
        for (int i = 0; i < MT; ++i)
#pragma unroll
          for (int j = 0; j < NT; ++j) dmma884(acc[i][j][0], acc[i][j][1], af[i], bf[j]);
      } else {
        // conjugation (sA, sB = -1) enters only through the sums and the epilogue: one FMA per fragment is all
        // the non-tensor FP64 work of a k-step (it shares the FP64 pipe with the DMMAs)
        double as_[MT], bs_[NT];
#pragma unroll
        for (int i = 0; i < MT; ++i) as_[i] = fma(sgnA, af[i].y, af[i].x);
#pragma unroll
        for (int j = 0; j < NT; ++j) bs_[j] = fma(sgnB, bf[j].y, bf[j].x);
#pragma unroll
        for (int i = 0; i < MT; ++i)
#pragma unroll
          for (int j = 0; j < NT; ++j) {
            dmma884(acc[i][j][0], acc[i][j][1], af[i].x, bf[j].x);  // P1 += ar*br
            dmma884(acc[i][j][2], acc[i][j][3], af[i].y, bf[j].y);  // P2 += ai*bi
            dmma884(acc[i][j][4], acc[i][j][5], as_[i], bs_[j]);    // P3 += (ar + sA ai)*(br + sB bi)
          }
      }
    }
  }
  cp_async_wait<0>();

  // epilogue: thread owns C[row = gq][cols 2*tq, 2*tq+1] of every 8x8 tile
  const bool has_beta = (g.beta_r != 0.0 || g.beta_i != 0.0) && g.splits <= 1;
  const double alpha_r = g.splits > 1 ? 1.0 : g.alpha_r, alpha_i = g.splits > 1 ? 0.0 : g.alpha_i;
#pragma unroll
  for (int i = 0; i < MT; ++i) {
    const int64_t row = m0 + wm * WM + i * 8 + gq;
    if (row >= g.M) continue;
#pragma unroll
    for (int j = 0; j < NT; ++j) {
      const int64_t col = n0 + wn * WN + j * 8 + 2 * tq;
      if (col >= g.N) continue;
      T* dst = Cg + row * ldc + col;
      if constexpr (!CPLX) {
        double v0 = alpha_r * acc[i][j][0], v1 = alpha_r * acc[i][j][1];
        if (has_beta) {
          v0 += g.beta_r * dst[0];
          if (col + 1 < g.N) v1 += g.beta_r * dst[1];
        }
        if (col + 1 < g.N && ((reinterpret_cast<uintptr_t>(dst) & 15) == 0)) {
          *reinterpret_cast<double2*>(dst) = make_double2(v0, v1);
        } else {
          dst[0] = v0;
          if (col + 1 < g.N) dst[1] = v1;
        }
      } else {
        const double2 al = make_double2(alpha_r, alpha_i), be = make_double2(g.beta_r, g.beta_i);
        const double p20 = sgnA * sgnB * acc[i][j][2], p21 = sgnA * sgnB * acc[i][j][3];
        double2 v0 = cmul(al, make_double2(acc[i][j][0] - p20, acc[i][j][4] - acc[i][j][0] - p20));
        double2 v1 = cmul(al, make_double2(acc[i][j][1] - p21, acc[i][j][5] - acc[i][j][1] - p21));
        if (has_beta) {
          v0 = cadd(v0, cmul(be, dst[0]));
          if (col + 1 < g.N) v1 = cadd(v1, cmul(be, dst[1]));
        }
        dst[0] = v0;
        if (col + 1 < g.N) dst[1] = v1;
      }
    }
  }
}

template <bool CPLX, bool SMALL, bool A_KC, bool B_KC, bool VEC>
static int launch_gemm(GemmArgs& g, int64_t batch, cudaStream_t st) {
  typedef Cfg<CPLX, SMALL> C_;
  typedef TileLayout<CPLX, A_KC, C_::BM, C_::BK> LA;
  typedef TileLayout<CPLX, B_KC, C_::BN, C_::BK> LB;
  constexpr size_t ES = CPLX ? 16 : 8;
  constexpr size_t smem = (size_t)STAGES * (LA::ELEMS + LB::ELEMS) * ES;
  auto kern = gemm_kernel<CPLX, SMALL, A_KC, B_KC, VEC>;
  static PerDeviceOnce once;
  if (once.need()) {
    TNB_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    once.done();
  }
  g.tiles_m = (int)((g.M + C_::BM - 1) / C_::BM);
  g.tiles_n = (int)((g.N + C_::BN - 1) / C_::BN);
  const int threads = C_::WARPS_M * C_::WARPS_N * 32;
  // grid.y carries the batch (<= 65535 per launch)
  for (int64_t b0 = 0; b0 < batch; b0 += 65535) {
    const int64_t nb = (batch - b0 < 65535) ? (batch - b0) : 65535;
    GemmArgs h = g;
    h.A = (const char*)g.A + b0 * g.sA * ES;
    h.B = (const char*)g.B + b0 * g.sB * ES;
    h.C = (char*)g.C + b0 * g.sC * ES;
    dim3 grid((unsigned)(g.tiles_m * g.tiles_n), (unsigned)nb, (unsigned)(g.splits > 1 ? g.splits : 1));
    TNB_CUDA_CHECK(launch_k(kern, grid, dim3(threads), smem, st, h));
    TNB_LAUNCH_CHECK();
  }
  return 0;
}

// ---- TMA-fed variant of the full-size complex tile ---------------------------------------------------
// Shared memory per stage: A tile 128 x 8 complex (16 KB) | B tile 64 x 8 complex (8 KB), dense, as the TMA
// boxes land with CU_TENSOR_MAP_SWIZZLE_128B: inside every 1 KB atom (8 rows of 128 bytes) the 16-byte
// chunk index is XORed with the row index.
//   K-contiguous operand  ([mn][k] in global): one box {8 k, BMN rows}: row = mn, chunk = k
//   MN-contiguous operand ([k][mn] in global): BMN / 8 boxes {8 mn, 8 k rows} of 1 KB: box = mn / 8,
//                                              row = k, chunk = mn % 8
// A DMMA fragment load is one 16-byte element per lane, row/column index gq = lane / 4, k index tq = lane % 4,
// served a quarter-warp (gq in {2i, 2i+1}, tq = 0..3) at a time.  In both layouts the chunk a lane reads is
// (index of gq's row inside its group of 8) XOR (k inside the stage); with the natural assignment the two rows of
// a quarter-warp differ in bit 0 only and collide on the same four chunks.  So fragment row gq is mapped to
// tile row sigma8(gq) = 4 (gq & 1) + (gq >> 1) inside its group of 8 -- the two rows then differ in bit 2 and
// the eight lanes cover all eight chunks.  The epilogue writes C through the same permutation.
constexpr int TMA_STAGES = 6;
constexpr int TMA_A_BYTES = 128 * 8 * 16, TMA_B_BYTES = 64 * 8 * 16, TMA_STAGE_BYTES = TMA_A_BYTES + TMA_B_BYTES;
__host__ __device__ __forceinline__ int sigma8(int g) { return ((g & 1) << 2) | (g >> 1); }

__device__ __forceinline__ void tma_load_3d(void* smem_dst, const CUtensorMap* map, int c0, int c1, int c2, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];\n" ::"r"(
          smem_u32(smem_dst)),
      "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
// bounded wait: a tensor map the hardware rejects would otherwise hang the grid
__device__ __forceinline__ void mbar_wait_bounded(uint64_t* bar, uint32_t parity) {
  for (uint32_t spin = 0; !mbar_try_wait(bar, parity); ++spin)
    if (spin > (1u << 26)) __trap();
}

template <bool A_KC, bool B_KC>
__global__ void __launch_bounds__(256) gemm_tma_kernel(GemmArgs g, const __grid_constant__ CUtensorMap mapA,
                                                       const __grid_constant__ CUtensorMap mapB) {
  typedef double2 T;
  constexpr int BM = 128, BN = 64, BK = 8, WM = 32, WN = 32, WARPS_N = 2;
  constexpr int MT = WM / 8, NT = WN / 8;
  extern __shared__ __align__(16) unsigned char smem_dyn[];
  // the swizzle atoms are 1 KB: align the stage buffers in the shared-memory window (the launch adds 1 KB of slack)
  unsigned char* smem_raw = smem_dyn + ((1024u - (smem_u32(smem_dyn) & 1023u)) & 1023u);
  __shared__ __align__(8) uint64_t full[TMA_STAGES];

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int gq = lane >> 2, tq = lane & 3, sg = sigma8(gq);
  const int wm = warp / WARPS_N, wn = warp % WARPS_N;
  const int tile = blockIdx.x;
  const int tm = tile % g.tiles_m, tn = tile / g.tiles_m;
  const int64_t m0 = (int64_t)tm * BM, n0 = (int64_t)tn * BN;
  const int bz = (int)blockIdx.y;
  T* Cg = reinterpret_cast<T*>(g.C) + (int64_t)bz * g.sC;
  const int64_t kbeg = (int64_t)blockIdx.z * g.k_per_split;
  const int64_t kend = (g.splits > 1 && kbeg + g.k_per_split < g.K) ? kbeg + g.k_per_split : g.K;
  int64_t ldc = g.ldc;
  if (g.splits > 1) {
    Cg = reinterpret_cast<T*>(g.part) + (int64_t)blockIdx.z * g.M * g.N;
    ldc = g.N;
  }
  if (tid == 0) {
#pragma unroll
    for (int s = 0; s < TMA_STAGES; ++s) mbar_init(&full[s], 1);
    fence_mbar_init();
  }
  __syncthreads();
  griddep_wait();               // barrier set-up above overlaps the tail of the previous kernel of a chain
  griddep_launch_dependents();
  if (g.skip && *g.skip) return;

  double acc[MT][NT][6];
#pragma unroll
  for (int i = 0; i < MT; ++i)
#pragma unroll
    for (int j = 0; j < NT; ++j)
#pragma unroll
      for (int r = 0; r < 6; ++r) acc[i][j][r] = 0.0;

  const int KT = (int)((kend - kbeg + BK - 1) / BK);   // split boundaries are multiples of BK; the tail is zero-filled
  auto issue = [&](int kt) {   // thread 0 only
    const int s = kt % TMA_STAGES;
    unsigned char* a_s = smem_raw + s * TMA_STAGE_BYTES;
    unsigned char* b_s = a_s + TMA_A_BYTES;
    const int k0 = (int)(kbeg + (int64_t)kt * BK);
    mbar_arrive_expect_tx(&full[s], TMA_STAGE_BYTES);
    if constexpr (A_KC) {
      tma_load_3d(a_s, &mapA, 2 * k0, (int)m0, bz, &full[s]);
    } else {
#pragma unroll
      for (int j = 0; j < BM / 8; ++j) tma_load_3d(a_s + j * 1024, &mapA, 2 * ((int)m0 + 8 * j), k0, bz, &full[s]);
    }
    if constexpr (B_KC) {
      tma_load_3d(b_s, &mapB, 2 * k0, (int)n0, bz, &full[s]);
    } else {
#pragma unroll
      for (int j = 0; j < BN / 8; ++j) tma_load_3d(b_s + j * 1024, &mapB, 2 * ((int)n0 + 8 * j), k0, bz, &full[s]);
    }
  };
  if (tid == 0) {
    for (int s = 0; s < TMA_STAGES && s < KT; ++s) issue(s);
  }
  const double sgnA = g.conjA ? -1.0 : 1.0, sgnB = g.conjB ? -1.0 : 1.0;

  for (int kt = 0; kt < KT; ++kt) {
    const int s = kt % TMA_STAGES;
    mbar_wait_bounded(&full[s], (uint32_t)((kt / TMA_STAGES) & 1));
    const unsigned char* a_s = smem_raw + s * TMA_STAGE_BYTES;
    const unsigned char* b_s = a_s + TMA_A_BYTES;
#pragma unroll
    for (int kk = 0; kk < BK; kk += 4) {
      const int k = kk + tq, ch = ((k ^ sg) << 4);
      T af[MT], bf[NT];
#pragma unroll
      for (int i = 0; i < MT; ++i) {
        const int grp = wm * (WM / 8) + i;   // group of 8 tile rows
        const int off = A_KC ? ((grp * 8 + sg) * 128 + ch) : (grp * 1024 + k * 128 + ch);
        af[i] = *reinterpret_cast<const T*>(a_s + off);
      }
#pragma unroll
      for (int j = 0; j < NT; ++j) {
        const int grp = wn * (WN / 8) + j;
        const int off = B_KC ? ((grp * 8 + sg) * 128 + ch) : (grp * 1024 + k * 128 + ch);
        bf[j] = *reinterpret_cast<const T*>(b_s + off);
      }
      double as_[MT], bs_[NT];   // conjugation enters through the sums and the epilogue only (see gemm_kernel)
#pragma unroll
      for (int i = 0; i < MT; ++i) as_[i] = fma(sgnA, af[i].y, af[i].x);
#pragma unroll
      for (int j = 0; j < NT; ++j) bs_[j] = fma(sgnB, bf[j].y, bf[j].x);
#pragma unroll
      for (int i = 0; i < MT; ++i)
#pragma unroll
        for (int j = 0; j < NT; ++j) {
          dmma884(acc[i][j][0], acc[i][j][1], af[i].x, bf[j].x);
          dmma884(acc[i][j][2], acc[i][j][3], af[i].y, bf[j].y);
          dmma884(acc[i][j][4], acc[i][j][5], as_[i], bs_[j]);
        }
    }
    __syncthreads();   // every warp is done with stage s: it may be refilled
    if (tid == 0 && kt + TMA_STAGES < KT) issue(kt + TMA_STAGES);
  }

  // epilogue: the thread's accumulator pair of tile (i, j) is C[fragment row gq][fragment columns 2 tq, 2 tq + 1],
  // i.e. tile row sigma8(gq) and tile columns sigma8(2 tq) = tq, sigma8(2 tq + 1) = 4 + tq of the 8 x 8 block
  const bool has_beta = (g.beta_r != 0.0 || g.beta_i != 0.0) && g.splits <= 1;
  const double2 al = g.splits > 1 ? make_double2(1.0, 0.0) : make_double2(g.alpha_r, g.alpha_i);
  const double2 be = make_double2(g.beta_r, g.beta_i);
#pragma unroll
  for (int i = 0; i < MT; ++i) {
    const int64_t row = m0 + wm * WM + i * 8 + sg;
    if (row >= g.M) continue;
#pragma unroll
    for (int j = 0; j < NT; ++j) {
      const int64_t cb = n0 + wn * WN + j * 8;
      T* dst = Cg + row * ldc + cb;
      const double p20 = sgnA * sgnB * acc[i][j][2], p21 = sgnA * sgnB * acc[i][j][3];
      double2 v0 = cmul(al, make_double2(acc[i][j][0] - p20, acc[i][j][4] - acc[i][j][0] - p20));
      double2 v1 = cmul(al, make_double2(acc[i][j][1] - p21, acc[i][j][5] - acc[i][j][1] - p21));
      if (cb + tq < g.N) {
        if (has_beta) v0 = cadd(v0, cmul(be, dst[tq]));
        dst[tq] = v0;
      }
      if (cb + 4 + tq < g.N) {
        if (has_beta) v1 = cadd(v1, cmul(be, dst[4 + tq]));
        dst[4 + tq] = v1;
      }
    }
  }
}

// cuTensorMapEncodeTiled through the runtime's driver entry point (libtnb does not link libcuda)
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn encode_tiled_fn() {
  static std::atomic<void*> cached{nullptr};
  void* f = cached.load(std::memory_order_acquire);
  if (!f) {
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) != cudaSuccess ||
        q != cudaDriverEntryPointSuccess)
      f = nullptr;
    cudaGetLastError();
    cached.store(f, std::memory_order_release);
  }
  return (EncodeTiledFn)f;
}

// Tensor map of one complex128 operand viewed as float64 pairs: dims {2 * inner, outer, batch}.
// kc: global [mn][k] (inner = K, outer = MN, box {8 k, bmn rows});  !kc: global [k][mn] (inner = MN, outer = K,
// box {8 mn, 8 k rows}).  Returns false when the operand cannot be described (the caller uses the cp.async kernel).
static bool make_operand_map(CUtensorMap* map, const void* base, bool kc, int64_t MN, int64_t K, int64_t ld, int64_t stride,
                             int64_t batch, int bmn) {
  EncodeTiledFn enc = encode_tiled_fn();
  if (!enc) return false;
  const int64_t inner = kc ? K : MN, outer = kc ? MN : K;
  if (((uintptr_t)base & 15) != 0 || ld < inner || inner <= 0 || outer <= 0) return false;
  if (2 * inner >= (1LL << 31) || outer >= (1LL << 31) || batch >= (1LL << 31)) return false;
  if (batch > 1 && stride <= 0) return false;
  const cuuint64_t dims[3] = {(cuuint64_t)(2 * inner), (cuuint64_t)outer, (cuuint64_t)batch};
  const cuuint64_t row_bytes = (cuuint64_t)ld * 16;
  const cuuint64_t strides[2] = {row_bytes, batch > 1 ? (cuuint64_t)stride * 16 : row_bytes * (cuuint64_t)outer};
  if (strides[0] >= (1ULL << 40) || strides[1] >= (1ULL << 40)) return false;
  const cuuint32_t box[3] = {16, (cuuint32_t)(kc ? bmn : 8), 1};
  const cuuint32_t estr[3] = {1, 1, 1};
  return enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, const_cast<void*>(base), dims, strides, box, estr,
             CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

template <bool A_KC, bool B_KC>
static int launch_gemm_tma(GemmArgs& g, const CUtensorMap& ma, const CUtensorMap& mb, int64_t batch, cudaStream_t st) {
  constexpr size_t smem = (size_t)TMA_STAGES * TMA_STAGE_BYTES + 1024;   // + slack: the base is aligned up to 1 KB
  auto kern = gemm_tma_kernel<A_KC, B_KC>;
  static PerDeviceOnce once;
  if (once.need()) {
    TNB_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    once.done();
  }
  g.tiles_m = (int)((g.M + 127) / 128);
  g.tiles_n = (int)((g.N + 63) / 64);
  dim3 grid((unsigned)(g.tiles_m * g.tiles_n), (unsigned)batch, (unsigned)(g.splits > 1 ? g.splits : 1));
  TNB_CUDA_CHECK(launch_k(kern, grid, dim3(256), smem, st, g, ma, mb));
  TNB_LAUNCH_CHECK();
  return 0;
}

// The TMA path for the full-size complex tile; returns -1 when it does not apply.
static int try_gemm_tma(GemmArgs& g, bool a_kc, bool b_kc, int64_t batch, cudaStream_t st) {
  if (batch > 65535) return -1;
  CUtensorMap ma, mb;
  if (!make_operand_map(&ma, g.A, a_kc, g.M, g.K, g.lda, g.sA, batch, 128)) return -1;
  if (!make_operand_map(&mb, g.B, b_kc, g.N, g.K, g.ldb, g.sB, batch, 64)) return -1;
  if (a_kc && b_kc) return launch_gemm_tma<true, true>(g, ma, mb, batch, st);
  if (a_kc && !b_kc) return launch_gemm_tma<true, false>(g, ma, mb, batch, st);
  if (!a_kc && b_kc) return launch_gemm_tma<false, true>(g, ma, mb, batch, st);
  return launch_gemm_tma<false, false>(g, ma, mb, batch, st);
}

// ---- TMA-fed variant of the full-size float64 tile (128 x 128 x 16, 8 warps as 2 x 4, warp tile 64 x 32) -------------
// Stages of 16 KB + 16 KB, dense, 128-byte swizzle as above; an element is 8 bytes, so a fragment load is served a
// HALF-warp at a time (fragment rows gq = 0..3 or 4..7, k = kk + tq) and its 16 lanes must hit 16 different 8-byte slots:
//   K-contiguous  ([mn][k]): one box {16 k, 128 rows}; 16-byte chunk = (k >> 1) ^ (row & 7).  Fragment row gq -> tile
//     row sigma8r(gq) = 2 (gq & 3) + (gq >> 2): four rows whose low bits are 0 2 4 6 (or 1 3 5 7) spread the two chunks
//     of a k quadruple over all eight;
//   MN-contiguous ([k][mn]): a box {16 mn, 16 k} of 2 KB per group of 16 rows; chunk = ((mn & 15) >> 1) ^ (k & 7).
//     The fragment rows of TWO adjacent 8-row tiles map into one group of 16: pi16(tile parity, gq) =
//     4 parity + [0 1 8 9 2 3 10 11][gq] -- rows {0,1,8,9} give chunks {0..3} and {4..7} in both halves of the pair.
// The epilogue writes C through the same permutations (of the rows by A's layout, of the columns by B's).
constexpr int TMAR_STAGES = 5;
constexpr int TMAR_OP_BYTES = 128 * 16 * 8, TMAR_STAGE_BYTES = 2 * TMAR_OP_BYTES;
__host__ __device__ __forceinline__ int sigma8r(int g) { return ((g & 3) << 1) | (g >> 2); }
__host__ __device__ __forceinline__ int pi16(int parity, int g) { return (parity << 2) | ((g & 2) << 2) | ((g & 4) >> 1) | (g & 1); }
// row (or column) of fragment index g of 8-tile t inside a warp tile, by the operand's layout
template <bool KC> __device__ __forceinline__ int frag_pos(int t, int g) {
  if constexpr (KC) return t * 8 + sigma8r(g);
  else return (t >> 1) * 16 + pi16(t & 1, g);
}

template <bool A_KC, bool B_KC>
__global__ void __launch_bounds__(256) gemm_tma_real_kernel(GemmArgs g, const __grid_constant__ CUtensorMap mapA,
                                                            const __grid_constant__ CUtensorMap mapB) {
  constexpr int BM = 128, BN = 128, BK = 16, WM = 64, WN = 32, WARPS_N = 4;
  constexpr int MT = WM / 8, NT = WN / 8;
  extern __shared__ __align__(16) unsigned char smem_dyn[];
  unsigned char* smem_raw = smem_dyn + ((1024u - (smem_u32(smem_dyn) & 1023u)) & 1023u);
  __shared__ __align__(8) uint64_t full[TMAR_STAGES];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int gq = lane >> 2, tq = lane & 3;
  const int wm = warp / WARPS_N, wn = warp % WARPS_N;
  const int tile = blockIdx.x;
  const int tm = tile % g.tiles_m, tn = tile / g.tiles_m;
  const int64_t m0 = (int64_t)tm * BM, n0 = (int64_t)tn * BN;
  const int bz = (int)blockIdx.y;
  double* Cg = reinterpret_cast<double*>(g.C) + (int64_t)bz * g.sC;
  const int64_t kbeg = (int64_t)blockIdx.z * g.k_per_split;
  const int64_t kend = (g.splits > 1 && kbeg + g.k_per_split < g.K) ? kbeg + g.k_per_split : g.K;
  int64_t ldc = g.ldc;
  if (g.splits > 1) {
    Cg = reinterpret_cast<double*>(g.part) + (int64_t)blockIdx.z * g.M * g.N;
    ldc = g.N;
  }
  if (tid == 0) {
#pragma unroll
    for (int s = 0; s < TMAR_STAGES; ++s) mbar_init(&full[s], 1);
    fence_mbar_init();
  }
  __syncthreads();
  griddep_wait();
  griddep_launch_dependents();
  if (g.skip && *g.skip) return;

  double acc[MT][NT][2];
#pragma unroll
  for (int i = 0; i < MT; ++i)
#pragma unroll
    for (int j = 0; j < NT; ++j) { acc[i][j][0] = 0.0; acc[i][j][1] = 0.0; }

  const int KT = (int)((kend - kbeg + BK - 1) / BK);
  auto issue = [&](int kt) {   // thread 0 only
    const int s = kt % TMAR_STAGES;
    unsigned char* a_s = smem_raw + s * TMAR_STAGE_BYTES;
    unsigned char* b_s = a_s + TMAR_OP_BYTES;
    const int k0 = (int)(kbeg + (int64_t)kt * BK);
    mbar_arrive_expect_tx(&full[s], TMAR_STAGE_BYTES);
    if constexpr (A_KC) {
      tma_load_3d(a_s, &mapA, k0, (int)m0, bz, &full[s]);
    } else {
#pragma unroll
      for (int j = 0; j < BM / 16; ++j) tma_load_3d(a_s + j * 2048, &mapA, (int)m0 + 16 * j, k0, bz, &full[s]);
    }
    if constexpr (B_KC) {
      tma_load_3d(b_s, &mapB, k0, (int)n0, bz, &full[s]);
    } else {
#pragma unroll
      for (int j = 0; j < BN / 16; ++j) tma_load_3d(b_s + j * 2048, &mapB, (int)n0 + 16 * j, k0, bz, &full[s]);
    }
  };
  if (tid == 0) {
    for (int s = 0; s < TMAR_STAGES && s < KT; ++s) issue(s);
  }
  // byte offset of fragment element (tile t of the warp, fragment index gq, k) inside an operand buffer
  auto frag_off = [&](bool kc, int wbase, int t, int k) -> int {
    if (kc) {
      const int r = wbase + t * 8 + sigma8r(gq);
      return r * 128 + ((((k >> 1) ^ (r & 7))) << 4) + ((k & 1) << 3);
    }
    const int box = (wbase >> 4) + (t >> 1), p = pi16(t & 1, gq);
    return box * 2048 + k * 128 + ((((p >> 1) ^ (k & 7))) << 4) + ((p & 1) << 3);
  };
  for (int kt = 0; kt < KT; ++kt) {
    const int s = kt % TMAR_STAGES;
    mbar_wait_bounded(&full[s], (uint32_t)((kt / TMAR_STAGES) & 1));
    const unsigned char* a_s = smem_raw + s * TMAR_STAGE_BYTES;
    const unsigned char* b_s = a_s + TMAR_OP_BYTES;
#pragma unroll
    for (int kk = 0; kk < BK; kk += 4) {
      const int k = kk + tq;
      double af[MT], bf[NT];
#pragma unroll
      for (int i = 0; i < MT; ++i) af[i] = *reinterpret_cast<const double*>(a_s + frag_off(A_KC, wm * WM, i, k));
#pragma unroll
      for (int j = 0; j < NT; ++j) bf[j] = *reinterpret_cast<const double*>(b_s + frag_off(B_KC, wn * WN, j, k));
#pragma unroll
      for (int i = 0; i < MT; ++i)
#pragma unroll
        for (int j = 0; j < NT; ++j) dmma884(acc[i][j][0], acc[i][j][1], af[i], bf[j]);
    }
    __syncthreads();
    if (tid == 0 && kt + TMAR_STAGES < KT) issue(kt + TMAR_STAGES);
  }
  const bool has_beta = g.beta_r != 0.0 && g.splits <= 1;
  const double alpha = g.splits > 1 ? 1.0 : g.alpha_r;
#pragma unroll
  for (int i = 0; i < MT; ++i) {
    const int64_t row = m0 + wm * WM + frag_pos<A_KC>(i, gq);
    if (row >= g.M) continue;
#pragma unroll
    for (int j = 0; j < NT; ++j) {
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int64_t col = n0 + wn * WN + frag_pos<B_KC>(j, 2 * tq + e);
        if (col >= g.N) continue;
        double* dst = Cg + row * ldc + col;
        double v = alpha * acc[i][j][e];
        if (has_beta) v += g.beta_r * dst[0];
        dst[0] = v;
      }
    }
  }
}

static bool make_operand_map_real(CUtensorMap* map, const void* base, bool kc, int64_t MN, int64_t K, int64_t ld, int64_t stride,
                                  int64_t batch) {
  EncodeTiledFn enc = encode_tiled_fn();
  if (!enc) return false;
  const int64_t inner = kc ? K : MN, outer = kc ? MN : K;
  if (((uintptr_t)base & 15) != 0 || (ld & 1) != 0 || ld < inner || inner <= 0 || outer <= 0) return false;
  if (inner >= (1LL << 31) || outer >= (1LL << 31) || batch >= (1LL << 31)) return false;
  if (batch > 1 && (stride <= 0 || (stride & 1) != 0)) return false;
  const cuuint64_t dims[3] = {(cuuint64_t)inner, (cuuint64_t)outer, (cuuint64_t)batch};
  const cuuint64_t row_bytes = (cuuint64_t)ld * 8;
  const cuuint64_t strides[2] = {row_bytes, batch > 1 ? (cuuint64_t)stride * 8 : row_bytes * (cuuint64_t)outer};
  if (strides[0] >= (1ULL << 40) || strides[1] >= (1ULL << 40)) return false;
  const cuuint32_t box[3] = {16, (cuuint32_t)(kc ? 128 : 16), 1};
  const cuuint32_t estr[3] = {1, 1, 1};
  return enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, const_cast<void*>(base), dims, strides, box, estr,
             CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

template <bool A_KC, bool B_KC>
static int launch_gemm_tma_real(GemmArgs& g, const CUtensorMap& ma, const CUtensorMap& mb, int64_t batch, cudaStream_t st) {
  constexpr size_t smem = (size_t)TMAR_STAGES * TMAR_STAGE_BYTES + 1024;
  auto kern = gemm_tma_real_kernel<A_KC, B_KC>;
  static PerDeviceOnce once;
  if (once.need()) {
    TNB_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    once.done();
  }
  g.tiles_m = (int)((g.M + 127) / 128);
  g.tiles_n = (int)((g.N + 127) / 128);
  dim3 grid((unsigned)(g.tiles_m * g.tiles_n), (unsigned)batch, (unsigned)(g.splits > 1 ? g.splits : 1));
  TNB_CUDA_CHECK(launch_k(kern, grid, dim3(256), smem, st, g, ma, mb));
  TNB_LAUNCH_CHECK();
  return 0;
}

// The TMA path for the full-size float64 tile; returns -1 when it does not apply.
static int try_gemm_tma_real(GemmArgs& g, bool a_kc, bool b_kc, int64_t batch, cudaStream_t st) {
  if (batch > 65535) return -1;
  CUtensorMap ma, mb;
  if (!make_operand_map_real(&ma, g.A, a_kc, g.M, g.K, g.lda, g.sA, batch)) return -1;
  if (!make_operand_map_real(&mb, g.B, b_kc, g.N, g.K, g.ldb, g.sB, batch)) return -1;
  if (a_kc && b_kc) return launch_gemm_tma_real<true, true>(g, ma, mb, batch, st);
  if (a_kc && !b_kc) return launch_gemm_tma_real<true, false>(g, ma, mb, batch, st);
  if (!a_kc && b_kc) return launch_gemm_tma_real<false, true>(g, ma, mb, batch, st);
  return launch_gemm_tma_real<false, false>(g, ma, mb, batch, st);
}

// C = alpha * sum_z part[z] + beta * C  (deterministic: fixed summation order)
template <typename T>
__global__ void splitk_reduce_kernel(const T* part, int splits, int64_t M, int64_t N, T* C, int64_t ldc, double ar,
                                     double ai, double br, double bi, const int* skip) {
  griddep_wait();
  griddep_launch_dependents();
  if (skip && *skip) return;
  const int64_t step = (int64_t)gridDim.x * blockDim.x;
  const bool has_beta = (br != 0.0 || bi != 0.0);
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < M * N; i += step) {
    T s = part[i];
    for (int z = 1; z < splits; ++z) s = Num<T>::add(s, part[(int64_t)z * M * N + i]);
    const int64_t r = i / N, c = i - r * N;
    T& dst = C[r * ldc + c];
    if constexpr (sizeof(T) == 8) {
      dst = has_beta ? ar * s + br * dst : ar * s;
    } else {
      T v = cmul(make_double2(ar, ai), s);
      if (has_beta) v = cadd(v, cmul(make_double2(br, bi), dst));
      dst = v;
    }
  }
}

template <bool CPLX, bool SMALL, bool VEC>
static int dispatch_layout(GemmArgs& g, bool a_kc, bool b_kc, int64_t batch, cudaStream_t st) {
  if (a_kc && b_kc) return launch_gemm<CPLX, SMALL, true, true, VEC>(g, batch, st);
  if (a_kc && !b_kc) return launch_gemm<CPLX, SMALL, true, false, VEC>(g, batch, st);
  if (!a_kc && b_kc) return launch_gemm<CPLX, SMALL, false, true, VEC>(g, batch, st);
  return launch_gemm<CPLX, SMALL, false, false, VEC>(g, batch, st);
}

// C := beta*C when K == 0 or alpha == 0 (BLAS semantics)
template <typename T>
__global__ void scale_c_kernel(T* C, int64_t M, int64_t N, int64_t ldc, int64_t sC, double br, double bi) {
  T* c = C + (int64_t)blockIdx.y * sC;
  const int64_t step = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < M * N; i += step) {
    const int64_t r = i / N, col = i - r * N;
    T& x = c[r * ldc + col];
    if (br == 0.0 && bi == 0.0) x = Num<T>::zero();
    else if constexpr (sizeof(T) == 8) x = x * br;
    else x = cmul(x, make_double2(br, bi));
  }
}

int gemm_ws(int dtype, int opA, int opB, int64_t M, int64_t N, int64_t K, double ar, double ai, const void* A,
            int64_t lda, int64_t sA, const void* B, int64_t ldb, int64_t sB, double br, double bi, void* C, int64_t ldc,
            int64_t sC, int64_t batch, void* splitk_ws, size_t splitk_bytes, cudaStream_t st);

int gemm(int dtype, int opA, int opB, int64_t M, int64_t N, int64_t K, double ar, double ai, const void* A,
         int64_t lda, int64_t sA, const void* B, int64_t ldb, int64_t sB, double br, double bi, void* C, int64_t ldc,
         int64_t sC, int64_t batch, cudaStream_t st) {
  return gemm_ws(dtype, opA, opB, M, N, K, ar, ai, A, lda, sA, B, ldb, sB, br, bi, C, ldc, sC, batch, nullptr, 0, st);
}

// As gemm(); with a scratch buffer, problems that have few C tiles but a long K (the V^H A products of
// the blocked QR, Gram-like shapes) are split along K over grid.z and reduced in a second kernel.
int gemm_ws(int dtype, int opA, int opB, int64_t M, int64_t N, int64_t K, double ar, double ai, const void* A,
            int64_t lda, int64_t sA, const void* B, int64_t ldb, int64_t sB, double br, double bi, void* C, int64_t ldc,
            int64_t sC, int64_t batch, void* splitk_ws, size_t splitk_bytes, cudaStream_t st) {
  if (M < 0 || N < 0 || K < 0 || batch < 0) return TNB_E_ARG;
  if (M == 0 || N == 0 || batch == 0) return 0;
  if (!C) return TNB_E_ARG;
  const bool cplx = (dtype == TNB_C128);
  if (!cplx && dtype != TNB_F64) return TNB_E_ARG;
  if (!cplx && (ai != 0.0 || bi != 0.0)) return TNB_E_ARG;
  if (K == 0 || (ar == 0.0 && ai == 0.0)) {
    if (br == 1.0 && bi == 0.0) return 0;
    int64_t blocks = (M * N + 255) / 256;
    if (blocks > 1024) blocks = 1024;
    for (int64_t b0 = 0; b0 < batch; b0 += 65535) {
      const int64_t nb = (batch - b0 < 65535) ? (batch - b0) : 65535;
      dim3 grid((unsigned)blocks, (unsigned)nb);
      if (cplx) scale_c_kernel<double2><<<grid, 256, 0, st>>>((double2*)C + b0 * sC, M, N, ldc, sC, br, bi);
      else scale_c_kernel<double><<<grid, 256, 0, st>>>((double*)C + b0 * sC, M, N, ldc, sC, br, bi);
      TNB_LAUNCH_CHECK();
    }
    return 0;
  }
  if (!A || !B) return TNB_E_ARG;
  // (a product queued under a device-side skip flag may not execute: it is credited no work in the profile)
  ProfScope prof(KC_GEMM, st, g_skip_flag ? 0.0 : (cplx ? 8.0 : 2.0) * (double)M * (double)N * (double)K * (double)batch);
  GemmArgs g;
  g.A = A; g.B = B; g.C = C;
  g.M = M; g.N = N; g.K = K;
  g.lda = lda; g.ldb = ldb; g.ldc = ldc;
  g.sA = sA; g.sB = sB; g.sC = sC;
  g.alpha_r = ar; g.alpha_i = ai; g.beta_r = br; g.beta_i = bi;
  g.splits = 1; g.k_per_split = K; g.part = nullptr;
  g.skip = g_skip_flag;
  g.conjA = cplx && (opA == TNB_OP_C || opA == TNB_OP_J);
  g.conjB = cplx && (opB == TNB_OP_C || opB == TNB_OP_J);
  // op N on A: stored M x K (k contiguous). op T/C: stored K x M (m contiguous).
  const bool a_kc = (opA == TNB_OP_N || opA == TNB_OP_J);
  // op N on B: stored K x N (n contiguous). op T/C: stored N x K (k contiguous).
  const bool b_kc = !(opB == TNB_OP_N || opB == TNB_OP_J);
  // skinny problems: smaller tiles waste fewer DMMAs and fill more SMs -- unless split-K can fill the
  // machine with full-size tiles (8 warps per CTA, twice the operand reuse), which is then preferred
  const int64_t big_tiles = cplx ? ((M + 127) / 128) * ((N + 63) / 64) : ((M + 127) / 128) * ((N + 127) / 128);
  bool small = (M <= 64) || (N <= (cplx ? 32 : 64)) || (big_tiles * batch < (int64_t)sm_count());
  if (splitk_ws && batch == 1) {
    const int64_t bk = cplx ? 8 : 16;
    const int64_t max_by_k = K / (bk * 16);  // keep >= 16 k-steps per split
    const int64_t max_by_ws = (int64_t)(splitk_bytes / ((size_t)M * N * (cplx ? 16 : 8)));
    // CTAs that are resident at once: the full-size tile needs 145 KB / 244 registers (one CTA per SM), the small
    // one fits twice.  The split aims at ONE full wave, rounded down (a few CTAs more than a wave cost a whole extra
    // wave: measured 89 -> 67 us on the 1152 x 64 x 3072 product of the QR, 357 -> 271 us on 1536 x 256 x 3072)
    auto plan = [&](int64_t tiles, bool small_tile) {
      const int64_t capacity = (int64_t)sm_count() * (small_tile ? 2 : 1);
      int64_t want = capacity / tiles;
      if (want < 1 && !small_tile && tiles < 2 * capacity) {
        // between one and two waves of full-size tiles (192 tiles on 148 SMs run at 65 %): a 2-4 way split that
        // lands just under a whole number of waves buys that back; the partials stay small next to the product
        double best = (double)tiles / (double)(((tiles + capacity - 1) / capacity) * capacity) + 0.15;
        for (int64_t s_ = 2; s_ <= 4; ++s_) {
          const double eff = (double)(tiles * s_) / (double)(((tiles * s_ + capacity - 1) / capacity) * capacity);
          if (eff > best) { best = eff; want = s_; }
        }
      }
      if (want > max_by_k) want = max_by_k;
      if (want > max_by_ws) want = max_by_ws;
      if (want > 64) want = 64;
      return want < 1 ? (int64_t)1 : want;
    };
    if (small && M > 64 && N > (cplx ? 32 : 64) && big_tiles * plan(big_tiles, false) * 5 >= (int64_t)sm_count() * 4) small = false;
    const int64_t bm = small ? 64 : 128, bn = cplx ? (small ? 32 : 64) : (small ? 64 : 128);
    const int64_t tiles = ((M + bm - 1) / bm) * ((N + bn - 1) / bn);
    const int64_t want = plan(tiles, small);
    if (want >= 2) {
      int64_t kps = (K + want - 1) / want;
      kps = ((kps + bk - 1) / bk) * bk;
      g.splits = (int)((K + kps - 1) / kps);
      g.k_per_split = kps;
      g.part = splitk_ws;
    }
  }
  auto finish = [&](int rc) -> int {
    if (rc != 0 || g.splits <= 1) return rc;
    int64_t blocks = (M * N + 255) / 256;
    if (blocks > (int64_t)sm_count() * 8) blocks = (int64_t)sm_count() * 8;
    if (cplx) TNB_CUDA_CHECK(launch_k(splitk_reduce_kernel<double2>, dim3((unsigned)blocks), dim3(256), 0, st, (const double2*)g.part, g.splits, M, N, (double2*)C, ldc, ar, ai, br, bi, g.skip));
    else TNB_CUDA_CHECK(launch_k(splitk_reduce_kernel<double>, dim3((unsigned)blocks), dim3(256), 0, st, (const double*)g.part, g.splits, M, N, (double*)C, ldc, ar, ai, br, bi, g.skip));
    TNB_LAUNCH_CHECK();
    return 0;
  };
  if (cplx) {
#ifndef TNB_EXP_NO_TMA
    if (!small) {
#else
    if (false) {   // kernel experiments: cp.async staging everywhere
#endif
      const int rt = try_gemm_tma(g, a_kc, b_kc, batch, st);
      if (rt >= 0) return finish(rt);
    }
    return finish(small ? dispatch_layout<true, true, true>(g, a_kc, b_kc, batch, st)
                        : dispatch_layout<true, false, true>(g, a_kc, b_kc, batch, st));
  }
#ifndef TNB_EXP_NO_TMA
  if (!small) {
    const int rt = try_gemm_tma_real(g, a_kc, b_kc, batch, st);
    if (rt >= 0) return finish(rt);
  }
#endif
  const bool vec = ((uintptr_t)A % 16 == 0) && ((uintptr_t)B % 16 == 0) && (lda % 2 == 0) && (ldb % 2 == 0) &&
                   (sA % 2 == 0 || batch == 1) && (sB % 2 == 0 || batch == 1);
  if (vec)
    return finish(small ? dispatch_layout<false, true, true>(g, a_kc, b_kc, batch, st)
                        : dispatch_layout<false, false, true>(g, a_kc, b_kc, batch, st));
  return finish(small ? dispatch_layout<false, true, false>(g, a_kc, b_kc, batch, st)
                      : dispatch_layout<false, false, false>(g, a_kc, b_kc, batch, st));
}

}  // namespace tnb

extern "C" int tnb_gemm(int dtype, int opA, int opB, int64_t M, int64_t N, int64_t K, const double* alpha,
                        const void* A, int64_t lda, int64_t strideA, const void* B, int64_t ldb, int64_t strideB,
                        const double* beta, void* C, int64_t ldc, int64_t strideC, int64_t batch, void* stream) {
  if (!alpha || !beta) return TNB_E_ARG;
  if (opA < 0 || opA > 3 || opB < 0 || opB > 3) return TNB_E_ARG;
  return tnb::gemm(dtype, opA, opB, M, N, K, alpha[0], alpha[1], A, lda, strideA, B, ldb, strideB, beta[0], beta[1],
                   C, ldc, strideC, batch, (cudaStream_t)stream);
}

extern "C" int tnb_gemm_ws(int dtype, int opA, int opB, int64_t M, int64_t N, int64_t K, const double* alpha,
                           const void* A, int64_t lda, const void* B, int64_t ldb, const double* beta, void* C,
                           int64_t ldc, void* ws, size_t ws_bytes, void* stream) {
  if (!alpha || !beta) return TNB_E_ARG;
  if (opA < 0 || opA > 3 || opB < 0 || opB > 3) return TNB_E_ARG;
  return tnb::gemm_ws(dtype, opA, opB, M, N, K, alpha[0], alpha[1], A, lda, 0, B, ldb, 0, beta[0], beta[1], C, ldc, 0, 1,
                      ws_bytes ? ws : nullptr, ws_bytes, (cudaStream_t)stream);
}
