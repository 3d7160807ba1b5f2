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
// one output row (l, a, p) and streams the columns (r, b) with coalesced
// stores; the A row it needs is read once per CTA (re-use across b via L1).
#include "common.cuh"

namespace tnb {

struct ApplyParams {
  int64_t d, Dl, Dr, wl, wr, dout;
  int64_t sAq, sAl, sAr;
  int64_t sWa, sWb, sWp, sWq;
};

template <typename T>
__global__ void __launch_bounds__(256) mps_mpo_site_kernel(const T* __restrict__ A, const T* __restrict__ W,
                                                           T* __restrict__ out, ApplyParams p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  T* Ws = reinterpret_cast<T*>(smem_raw);  // [b][q] slice of W for this CTA's (a, p)
  const int64_t row = blockIdx.x;           // (l, a, p)
  const int64_t pp = row % p.dout;
  const int64_t la = row / p.dout;
  const int64_t a = la % p.wl, l = la / p.wl;
  for (int i = threadIdx.x; i < p.wr * p.d; i += blockDim.x) {
    const int64_t b = i / p.d, q = i - b * p.d;
    Ws[i] = W[a * p.sWa + b * p.sWb + pp * p.sWp + q * p.sWq];
  }
  __syncthreads();
  const int64_t ncol = p.Dr * p.wr;
  const T* Arow = A + l * p.sAl;
  T* orow = out + row * ncol;
  const uint32_t wr = (uint32_t)p.wr;
  for (int64_t c0 = (int64_t)blockIdx.y * blockDim.x; c0 < ncol; c0 += (int64_t)gridDim.y * blockDim.x) {
    const int64_t col = c0 + threadIdx.x;
    if (col >= ncol) break;
    const uint32_t r = (uint32_t)col / wr, b = (uint32_t)col - r * wr;
    T acc = Num<T>::zero();
    for (int64_t q = 0; q < p.d; ++q) acc = Num<T>::fma(Arow[q * p.sAq + (int64_t)r * p.sAr], Ws[b * p.d + q], acc);
    orow[col] = acc;
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
  const int64_t rows = p.Dl * p.wl * p.dout, ncol = p.Dr * p.wr;
  if (rows == 0 || ncol == 0) return 0;
  if (p.d == 0) return TNB_E_ARG;
  if (rows > 2147483647LL || ncol > 4294967295LL) return TNB_E_UNSUPPORTED;
  const size_t smem = (size_t)(p.wr * p.d) * elem_size(A->dtype);
  if (smem > 48 * 1024) return TNB_E_UNSUPPORTED;
  // enough column slabs to fill the machine when there are few rows
  int64_t gy = 1;
  const int64_t want = (int64_t)sm_count() * 8;
  if (rows < want) gy = (want + rows - 1) / rows;
  const int64_t max_gy = (ncol + 255) / 256;
  if (gy > max_gy) gy = max_gy;
  if (gy > 65535) gy = 65535;
  dim3 grid((unsigned)rows, (unsigned)gy);
  cudaStream_t st = (cudaStream_t)stream;
  ProfScope prof(KC_MPO_APPLY, st, (double)elem_size(A->dtype) *
                                       ((double)rows * ncol + (double)numel(A) + (double)numel(W)));
  if (A->dtype == TNB_F64)
    mps_mpo_site_kernel<double><<<grid, 256, smem, st>>>((const double*)A->ptr, (const double*)W->ptr, (double*)out, p);
  else
    mps_mpo_site_kernel<double2><<<grid, 256, smem, st>>>((const double2*)A->ptr, (const double2*)W->ptr, (double2*)out, p);
  TNB_LAUNCH_CHECK();
  return 0;
}
