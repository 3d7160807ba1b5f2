// MPO x MPS site apply, fused with consolidate_indices()
// (reference: onedim_core.py:1702-1704 -> tensor.py:735 + tensor.py:340-370).
//
//   out[(l, a), p, (r, b)] = sum_q A[q, l, r] * W[a, b, p, q]
//
// K = d (2..16) is far too small for the tensor pipe to matter: the kernel is
// bound by writing `out` once (algorithmic bytes = |out| + |A| + |W|; for the
// chi=512, D=3, d=2 complex128 site that is 75.5 MB + 8.4 MB).  The reference
// path writes the K=d GEMM result and then permutes it -- three passes over
// the 75 MB tensor instead of one.  W lives in shared memory; one CTA handles
// the `dout` consecutive output rows (l, a, :) and streams the columns (r, b)
// with coalesced stores; each A value is loaded once for all p (re-use across
// b via L1).
#include "common.cuh"

namespace tnb {

struct ApplyParams {
  int64_t d, Dl, Dr, wl, wr, dout;
  int64_t pc;  // output rows p per CTA (blockIdx.z owns p in [z * pc, (z + 1) * pc)): their W slices share the 48 KB
  int64_t sAq, sAl, sAr;
  int64_t sWa, sWb, sWp, sWq;
};

// One CTA per (l, a) and column slab: it writes the `dout` consecutive output rows (l, a, p = 0..dout-1), so
// every A value is loaded once for all p, and walks its columns without integer division (r, b advance by
// fixed increments); the column loop is unrolled so that several independent loads are in flight per thread.
template <typename T, int DQ>   // DQ = d when d <= 4 (A values and products fully in registers), 0 = any d
__global__ void __launch_bounds__(256) mps_mpo_site_kernel(const T* __restrict__ A, const T* __restrict__ W,
                                                           T* __restrict__ out, ApplyParams p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  T* Ws = reinterpret_cast<T*>(smem_raw);  // [p][b][q] slice of W for this CTA's a
  const int64_t la = blockIdx.x;            // (l, a)
  const int64_t a = la % p.wl, l = la / p.wl;
  const int d = DQ ? DQ : (int)p.d;
  const int wrd = (int)p.wr * d;
  const int p0 = (int)(blockIdx.z * p.pc);
  const int np = (p.dout - p0 < p.pc) ? (int)(p.dout - p0) : (int)p.pc;
  for (int i = threadIdx.x; i < np * wrd; i += blockDim.x) {
    const int pp = i / wrd, rem = i - pp * wrd;
    const int b = rem / d, q = rem - b * d;
    Ws[i] = W[a * p.sWa + b * p.sWb + (p0 + pp) * p.sWp + q * p.sWq];
  }
  __syncthreads();
  const uint32_t ncol = (uint32_t)(p.Dr * p.wr), wr = (uint32_t)p.wr;
  const T* Arow = A + l * p.sAl;
  T* orow = out + (la * p.dout + p0) * (int64_t)ncol;
  const uint32_t step = blockDim.x * gridDim.y;
  const uint32_t dr = step / wr, db = step - dr * wr;
  uint32_t col = blockIdx.y * blockDim.x + threadIdx.x;
  uint32_t r = col / wr, b = col - r * wr;
#pragma unroll 2
  for (; col < ncol; col += step) {
    if constexpr (DQ > 0) {
      T av[DQ];
#pragma unroll
      for (int q = 0; q < DQ; ++q) av[q] = Arow[q * p.sAq + (int64_t)r * p.sAr];
      for (int pp = 0; pp < np; ++pp) {
        const T* w = Ws + pp * wrd + b * DQ;
        T acc = Num<T>::zero();
#pragma unroll
        for (int q = 0; q < DQ; ++q) acc = Num<T>::fma(av[q], w[q], acc);
        orow[(int64_t)pp * ncol + col] = acc;
      }
    } else {
      for (int pp = 0; pp < np; ++pp) {
        const T* w = Ws + pp * wrd + b * d;
        T acc = Num<T>::zero();
        for (int q = 0; q < d; ++q) acc = Num<T>::fma(Arow[q * p.sAq + (int64_t)r * p.sAr], w[q], acc);
        orow[(int64_t)pp * ncol + col] = acc;
      }
    }
    r += dr; b += db;
    if (b >= wr) { b -= wr; ++r; }
  }
}

}  // namespace tnb

using namespace tnb;

extern "C" int tnb_mps_mpo_site(const tnb_tensor_t* A, const tnb_tensor_t* W, void* out, void* stream) {
  if (!valid_tensor(A) || !valid_tensor(W) || !out || A->rank != 3 || W->rank != 4 || A->dtype != W->dtype)
    return TNB_E_ARG;
  if (A->shape[0] != W->shape[3]) return TNB_E_ARG;
  ApplyParams p;
  p.d = A->shape[0]; p.Dl = A->shape[1]; p.Dr = A->shape[2];
  p.wl = W->shape[0]; p.wr = W->shape[1]; p.dout = W->shape[2];
  p.sAq = A->stride[0]; p.sAl = A->stride[1]; p.sAr = A->stride[2];
  p.sWa = W->stride[0]; p.sWb = W->stride[1]; p.sWp = W->stride[2]; p.sWq = W->stride[3];
  const int64_t groups = p.Dl * p.wl, rows = groups * p.dout, ncol = p.Dr * p.wr;
  if (rows == 0 || ncol == 0) return 0;
  if (p.d == 0) return TNB_E_ARG;
  if (groups > 2147483647LL || ncol > 4294967295LL - 65536LL * 256) return TNB_E_UNSUPPORTED;
  const size_t slice = (size_t)(p.wr * p.d) * elem_size(A->dtype);   // W[a, :, p, :]
  if (slice > 48 * 1024) return TNB_E_UNSUPPORTED;
  p.pc = (int64_t)((48 * 1024) / slice);
  if (p.pc > p.dout) p.pc = p.dout;
  const int64_t gz = (p.dout + p.pc - 1) / p.pc;
  if (gz > 65535) return TNB_E_UNSUPPORTED;
  const size_t smem = (size_t)p.pc * slice;
  // enough column slabs to fill the machine when there are few (l, a) groups
  int64_t gy = 1;
  const int64_t want = (int64_t)sm_count() * 8;
  if (groups < want) gy = (want + groups - 1) / groups;
  const int64_t max_gy = (ncol + 255) / 256;
  if (gy > max_gy) gy = max_gy;
  if (gy > 65535) gy = 65535;
  dim3 grid((unsigned)groups, (unsigned)gy, (unsigned)gz);
  cudaStream_t st = (cudaStream_t)stream;
  ProfScope prof(KC_MPO_APPLY, st, (double)elem_size(A->dtype) *
                                       ((double)rows * ncol + (double)numel(A) + (double)numel(W)));
  const bool f64 = A->dtype == TNB_F64;
#define TNB_APPLY_LAUNCH(DQ)                                                                                          \
  do {                                                                                                                \
    if (f64) mps_mpo_site_kernel<double, DQ><<<grid, 256, smem, st>>>((const double*)A->ptr, (const double*)W->ptr, (double*)out, p); \
    else mps_mpo_site_kernel<double2, DQ><<<grid, 256, smem, st>>>((const double2*)A->ptr, (const double2*)W->ptr, (double2*)out, p); \
  } while (0)
  switch (p.d) {
    case 1: TNB_APPLY_LAUNCH(1); break;
    case 2: TNB_APPLY_LAUNCH(2); break;
    case 3: TNB_APPLY_LAUNCH(3); break;
    case 4: TNB_APPLY_LAUNCH(4); break;
    default: TNB_APPLY_LAUNCH(0); break;
  }
#undef TNB_APPLY_LAUNCH
  TNB_LAUNCH_CHECK();
  return 0;
}
