// Small elementwise / reduction kernels on device buffers: the direct
// ndarray arithmetic the reference's MPS layer does on Tensor.data
// (SURVEY.md section 8a row a14).  All HBM-bound, one pass over the data.
#include "common.cuh"

namespace tnb {

struct ViewParams {
  int rank;
  int64_t shape[TNB_MAX_RANK];
  int64_t stride[TNB_MAX_RANK];
  int64_t numel;
};

static ViewParams make_view(const tnb_tensor_t* t) {
  View v = collapse(t->rank, t->shape, t->stride);
  ViewParams p;
  p.rank = v.rank;
  p.numel = v.numel;
  for (int d = 0; d < v.rank; ++d) { p.shape[d] = v.shape[d]; p.stride[d] = v.stride[d]; }
  return p;
}

__device__ __forceinline__ int64_t view_offset(const ViewParams& p, int64_t idx) {
  int64_t off = 0;
#pragma unroll 1
  for (int d = p.rank - 1; d > 0; --d) {
    const int64_t s = p.shape[d];
    const int64_t q = idx / s;
    off += (idx - q * s) * p.stride[d];
    idx = q;
  }
  return off + idx * p.stride[0];
}

static inline unsigned grid_for(int64_t n, int threads = 256) {
  int64_t b = (n + threads - 1) / threads;
  const int64_t cap = (int64_t)sm_count() * 16;
  if (b > cap) b = cap;
  if (b < 1) b = 1;
  return (unsigned)b;
}

// ---- x *= alpha (strided view, in place) --------------------------------------
template <typename T>
__global__ void scale_inplace_kernel(T* x, ViewParams p, double ar, double ai) {
  const int64_t step = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < p.numel; i += step) {
    const int64_t off = view_offset(p, i);
    if constexpr (sizeof(T) == 8) x[off] = x[off] * ar;
    else x[off] = cmul(x[off], make_double2(ar, ai));
  }
}

// ---- out = a*x + b*y -----------------------------------------------------------
template <typename T>
__global__ void axpby_kernel(const T* x, ViewParams px, const T* y, ViewParams py, T* out, double ar, double ai,
                             double br, double bi) {
  const int64_t step = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < px.numel; i += step) {
    const T xv = x[view_offset(px, i)], yv = y[view_offset(py, i)];
    if constexpr (sizeof(T) == 8) out[i] = ar * xv + br * yv;
    else out[i] = cadd(cmul(xv, make_double2(ar, ai)), cmul(yv, make_double2(br, bi)));
  }
}

// ---- Frobenius norm --------------------------------------------------------------
// Deterministic two-level reduction: fixed grid, per-block partials in `ws`,
// the last block to finish (atomic ticket) sums them in index order.
constexpr int NORM_BLOCKS = 592;  // 4 x 148
constexpr int NORM_THREADS = 256;

template <typename T>
__global__ void __launch_bounds__(NORM_THREADS) norm2_kernel(const T* x, ViewParams p, double* out, double* partial,
                                                             unsigned int* ticket) {
  __shared__ double sh[NORM_THREADS / 32];
  __shared__ bool is_last;
  double acc = 0.0;
  const int64_t step = (int64_t)gridDim.x * blockDim.x;
  const bool contiguous = (p.rank == 1 && p.stride[0] == 1);
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < p.numel; i += step) {
    const T v = x[contiguous ? i : view_offset(p, i)];
    acc += Num<T>::abs2(v);
  }
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double s = 0.0;
    for (int w = 0; w < NORM_THREADS / 32; ++w) s += sh[w];
    partial[blockIdx.x] = s;
    __threadfence();
    const unsigned int t = atomicAdd(ticket, 1u);
    is_last = (t == gridDim.x - 1);
  }
  __syncthreads();
  if (is_last) {
    double s = 0.0;
    for (int i = threadIdx.x; i < (int)gridDim.x; i += blockDim.x) s += ((volatile double*)partial)[i];
    s = warp_sum(s);
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
      double tot = 0.0;
      for (int w = 0; w < NORM_THREADS / 32; ++w) tot += sh[w];
      *out = sqrt(tot);
      *ticket = 0u;  // ready for the next call on this workspace
    }
  }
}

// ---- diagonal helpers ---------------------------------------------------------------
__device__ __forceinline__ double diag_fn(double s, int mode) {
  return mode == 1 ? sqrt(s) : (mode == 2 ? 1.0 / s : s);
}

template <typename T>
__global__ void diag_embed_kernel(const double* s, int64_t n, T* out, int mode) {
  const int64_t step = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n * n; i += step) {
    const int64_t r = i / n, c = i - r * n;
    out[i] = (r == c) ? Num<T>::from(diag_fn(s[r], mode), 0.0) : Num<T>::zero();
  }
}

template <typename T>
__global__ void diag_extract_kernel(const T* x, int64_t n, int64_t s0, int64_t s1, T* out) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = x[i * (s0 + s1)];
}

template <typename T>
__global__ void diag_scale_kernel(T* x, int64_t rows, int64_t cols, int64_t ld, const double* s, int axis, int mode) {
  const int64_t step = (int64_t)gridDim.x * blockDim.x;
  const int64_t n = rows * cols;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += step) {
    const int64_t r = i / cols, c = i - r * cols;
    const double f = diag_fn(s[axis == 0 ? r : c], mode);
    x[r * ld + c] = Num<T>::scale(x[r * ld + c], f);
  }
}

// ---- trace over two axes ---------------------------------------------------------------
template <typename T>
__global__ void trace_kernel(const T* x, ViewParams rest, int64_t n, int64_t sdiag, T* out) {
  const int64_t step = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < rest.numel; i += step) {
    const int64_t base = view_offset(rest, i);
    T acc = Num<T>::zero();
    for (int64_t k = 0; k < n; ++k) acc = Num<T>::add(acc, x[base + k * sdiag]);
    out[i] = acc;
  }
}

__global__ void real_to_complex_kernel(const double* x, int64_t n, double2* out) {
  const int64_t step = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += step) out[i] = make_double2(x[i], 0.0);
}

// kept-rank rule, one CTA (n is a bond dimension: at most a few thousand)
__global__ void truncation_count_kernel(const double* s, int64_t n, int64_t chi, double threshold, int relative,
                                        double* info, double* s_scaled) {
  __shared__ int cnt;
  if (threadIdx.x == 0) cnt = 0;
  __syncthreads();
  const double s0 = n > 0 ? s[0] : 0.0;
  const double bar = relative ? threshold * s0 : threshold;
  const int64_t lim = (chi > 0 && chi < n) ? chi : n;
  int local = 0;
  for (int64_t i = threadIdx.x; i < n; i += blockDim.x) {
    const double v = s[i];
    // relative == 2: the comparison of onedim_core.py:333-336, s/s0 > threshold, bit for bit
    const bool above = (relative == 2) ? (v / s0 > threshold) : (v > bar);
    if (i < lim && above) local++;
    if (s_scaled) s_scaled[i] = v / s0;
  }
  atomicAdd(&cnt, local);
  __syncthreads();
  if (threadIdx.x == 0) { info[0] = (double)cnt; info[1] = s0; }
}

}  // namespace tnb

using namespace tnb;

extern "C" int tnb_scale_inplace(const tnb_tensor_t* x, double ar, double ai, void* stream) {
  if (!valid_tensor(x) || !x->ptr) return TNB_E_ARG;
  if (x->dtype == TNB_F64 && ai != 0.0) return TNB_E_ARG;
  ViewParams p = make_view(x);
  if (p.numel == 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  ProfScope prof(KC_ELEMWISE, st, 2.0 * (double)p.numel * (double)elem_size(x->dtype));
  if (x->dtype == TNB_F64) scale_inplace_kernel<double><<<grid_for(p.numel), 256, 0, st>>>((double*)x->ptr, p, ar, ai);
  else scale_inplace_kernel<double2><<<grid_for(p.numel), 256, 0, st>>>((double2*)x->ptr, p, ar, ai);
  TNB_LAUNCH_CHECK();
  return 0;
}

extern "C" int tnb_axpby(const tnb_tensor_t* x, const tnb_tensor_t* y, void* out, double ar, double ai, double br,
                         double bi, void* stream) {
  if (!valid_tensor(x) || !valid_tensor(y) || !out || x->dtype != y->dtype || x->rank != y->rank) return TNB_E_ARG;
  for (int i = 0; i < x->rank; ++i)
    if (x->shape[i] != y->shape[i]) return TNB_E_ARG;
  // do not collapse independently: both views must share one index space
  ViewParams px, py;
  px.rank = py.rank = x->rank > 0 ? x->rank : 1;
  px.numel = py.numel = numel(x);
  if (x->rank == 0) { px.shape[0] = py.shape[0] = 1; px.stride[0] = py.stride[0] = 1; }
  for (int i = 0; i < x->rank; ++i) {
    px.shape[i] = py.shape[i] = x->shape[i];
    px.stride[i] = x->stride[i];
    py.stride[i] = y->stride[i];
  }
  if (px.numel == 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  if (x->dtype == TNB_F64)
    axpby_kernel<double><<<grid_for(px.numel), 256, 0, st>>>((const double*)x->ptr, px, (const double*)y->ptr, py,
                                                            (double*)out, ar, ai, br, bi);
  else
    axpby_kernel<double2><<<grid_for(px.numel), 256, 0, st>>>((const double2*)x->ptr, px, (const double2*)y->ptr, py,
                                                             (double2*)out, ar, ai, br, bi);
  TNB_LAUNCH_CHECK();
  return 0;
}

extern "C" size_t tnb_norm2_workspace(void) { return (NORM_BLOCKS + 2) * sizeof(double); }

// ws layout: [0] ticket (must be zero on first use; the kernel re-zeroes it), [1..] partials
extern "C" int tnb_norm2(const tnb_tensor_t* x, double* out_device, void* ws, size_t ws_bytes, void* stream) {
  if (!valid_tensor(x) || !out_device || !ws) return TNB_E_ARG;
  if (ws_bytes < tnb_norm2_workspace()) return TNB_E_WORKSPACE;
  cudaStream_t st = (cudaStream_t)stream;
  ViewParams p = make_view(x);
  if (p.numel == 0) { TNB_CUDA_CHECK(cudaMemsetAsync(out_device, 0, sizeof(double), st)); return 0; }
  unsigned int* ticket = (unsigned int*)ws;
  double* partial = (double*)ws + 1;
  int64_t blocks = (p.numel + NORM_THREADS * 4 - 1) / (NORM_THREADS * 4);
  if (blocks > NORM_BLOCKS) blocks = NORM_BLOCKS;
  if (blocks < 1) blocks = 1;
  TNB_CUDA_CHECK(cudaMemsetAsync(ticket, 0, sizeof(unsigned int), st));
  ProfScope prof(KC_ELEMWISE, st, (double)p.numel * (double)elem_size(x->dtype));
  if (x->dtype == TNB_F64)
    norm2_kernel<double><<<(unsigned)blocks, NORM_THREADS, 0, st>>>((const double*)x->ptr, p, out_device, partial, ticket);
  else
    norm2_kernel<double2><<<(unsigned)blocks, NORM_THREADS, 0, st>>>((const double2*)x->ptr, p, out_device, partial, ticket);
  TNB_LAUNCH_CHECK();
  return 0;
}

extern "C" int tnb_diag_embed(int dtype, const double* s, int64_t n, void* out, int mode, void* stream) {
  if (!s || !out || n < 0 || mode < 0 || mode > 2) return TNB_E_ARG;
  if (n == 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == TNB_F64) diag_embed_kernel<double><<<grid_for(n * n), 256, 0, st>>>(s, n, (double*)out, mode);
  else if (dtype == TNB_C128) diag_embed_kernel<double2><<<grid_for(n * n), 256, 0, st>>>(s, n, (double2*)out, mode);
  else return TNB_E_ARG;
  TNB_LAUNCH_CHECK();
  return 0;
}

extern "C" int tnb_diag_extract(const tnb_tensor_t* x, void* out, void* stream) {
  if (!valid_tensor(x) || x->rank != 2 || !out) return TNB_E_ARG;
  const int64_t n = x->shape[0] < x->shape[1] ? x->shape[0] : x->shape[1];
  if (n == 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  const unsigned blocks = (unsigned)((n + 255) / 256);
  if (x->dtype == TNB_F64)
    diag_extract_kernel<double><<<blocks, 256, 0, st>>>((const double*)x->ptr, n, x->stride[0], x->stride[1], (double*)out);
  else
    diag_extract_kernel<double2><<<blocks, 256, 0, st>>>((const double2*)x->ptr, n, x->stride[0], x->stride[1], (double2*)out);
  TNB_LAUNCH_CHECK();
  return 0;
}

extern "C" int tnb_diag_scale(int dtype, void* x, int64_t rows, int64_t cols, int64_t ld, const double* s, int axis,
                              int mode, void* stream) {
  if (!x || !s || rows < 0 || cols < 0 || ld < cols || (axis != 0 && axis != 1) || mode < 0 || mode > 2) return TNB_E_ARG;
  if (rows * cols == 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  ProfScope prof(KC_ELEMWISE, st, 2.0 * (double)rows * (double)cols * (double)elem_size(dtype));
  if (dtype == TNB_F64) diag_scale_kernel<double><<<grid_for(rows * cols), 256, 0, st>>>((double*)x, rows, cols, ld, s, axis, mode);
  else if (dtype == TNB_C128) diag_scale_kernel<double2><<<grid_for(rows * cols), 256, 0, st>>>((double2*)x, rows, cols, ld, s, axis, mode);
  else return TNB_E_ARG;
  TNB_LAUNCH_CHECK();
  return 0;
}

extern "C" int tnb_trace(const tnb_tensor_t* x, int axis1, int axis2, void* out, void* stream) {
  if (!valid_tensor(x) || !out || axis1 == axis2 || axis1 < 0 || axis2 < 0 || axis1 >= x->rank || axis2 >= x->rank)
    return TNB_E_ARG;
  tnb_tensor_t rest = *x;
  rest.rank = 0;
  for (int i = 0; i < x->rank; ++i) {
    if (i == axis1 || i == axis2) continue;
    rest.shape[rest.rank] = x->shape[i];
    rest.stride[rest.rank] = x->stride[i];
    rest.rank++;
  }
  ViewParams p = make_view(&rest);
  if (p.numel == 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  // np.trace sums the main diagonal, of length min(n1, n2) when the two axes differ in size
  const int64_t n = x->shape[axis1] < x->shape[axis2] ? x->shape[axis1] : x->shape[axis2];
  const int64_t sd = x->stride[axis1] + x->stride[axis2];
  if (x->dtype == TNB_F64) trace_kernel<double><<<grid_for(p.numel), 256, 0, st>>>((const double*)x->ptr, p, n, sd, (double*)out);
  else trace_kernel<double2><<<grid_for(p.numel), 256, 0, st>>>((const double2*)x->ptr, p, n, sd, (double2*)out);
  TNB_LAUNCH_CHECK();
  return 0;
}

extern "C" int tnb_real_to_complex(const double* x, int64_t n, void* out, void* stream) {
  if (!x || !out || n < 0) return TNB_E_ARG;
  if (n == 0) return 0;
  real_to_complex_kernel<<<grid_for(n), 256, 0, (cudaStream_t)stream>>>(x, n, (double2*)out);
  TNB_LAUNCH_CHECK();
  return 0;
}

extern "C" int tnb_truncation_count(const double* s, int64_t n, int64_t chi, double threshold, int relative,
                                    double* info_device, double* s_scaled, void* stream) {
  if (!s || !info_device || n < 0) return TNB_E_ARG;
  truncation_count_kernel<<<1, 256, 0, (cudaStream_t)stream>>>(s, n, chi, threshold, relative, info_device, s_scaled);
  TNB_LAUNCH_CHECK();
  return 0;
}

extern "C" int tnb_version(void) { return 100; }
extern "C" long long tnb_launch_count(int reset) {
  const long long v = reset ? tnb::g_launches.exchange(0) : tnb::g_launches.load();
  return v;
}

extern "C" const char* tnb_error_string(int code) {
  switch (code) {
    case TNB_OK: return "ok";
    case TNB_E_ARG: return "invalid argument";
    case TNB_E_WORKSPACE: return "workspace too small";
    case TNB_E_UNSUPPORTED: return "unsupported configuration";
    case TNB_E_NOCONV: return "Jacobi SVD did not converge";
    default: return code > 0 ? cudaGetErrorString((cudaError_t)code) : "unknown error";
  }
}

extern "C" int tnb_device_info(int* sms, int* major, int* minor) {
  int dev = 0;
  TNB_CUDA_CHECK(cudaGetDevice(&dev));
  if (sms) TNB_CUDA_CHECK(cudaDeviceGetAttribute(sms, cudaDevAttrMultiProcessorCount, dev));
  if (major) TNB_CUDA_CHECK(cudaDeviceGetAttribute(major, cudaDevAttrComputeCapabilityMajor, dev));
  if (minor) TNB_CUDA_CHECK(cudaDeviceGetAttribute(minor, cudaDevAttrComputeCapabilityMinor, dev));
  return 0;
}

// ---- counter-based uniform fill (batched-path input generation on the owning GPU) ------------
// out[i] = U[0,1) from splitmix64(key + (offset + i) * golden): any element can be regenerated on the
// host (tests/, batch.py:host_uniform) without replaying a stream.
namespace tnb {
__global__ void fill_uniform_kernel(double* out, int64_t n, unsigned long long key, unsigned long long offset) {
  const int64_t step = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += step) {
    unsigned long long z = key + (offset + (unsigned long long)i) * 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    z = z ^ (z >> 31);
    out[i] = (double)(z >> 11) * (1.0 / 9007199254740992.0);
  }
}
}  // namespace tnb

extern "C" int tnb_fill_uniform(double* out, int64_t n, unsigned long long key, unsigned long long offset,
                                void* stream) {
  if (!out || n < 0) return TNB_E_ARG;
  if (n == 0) return 0;
  tnb::fill_uniform_kernel<<<grid_for(n), 256, 0, (cudaStream_t)stream>>>(out, n, key, offset);
  TNB_LAUNCH_CHECK();
  return 0;
}

// ---- batched helpers of the batched path (one launch for the same site of every network of a shard) --------
namespace tnb {
// matrix b: out[b * per + i] = U[0,1) from stream keys[b]
__global__ void fill_uniform_batched_kernel(double* out, int64_t per, const unsigned long long* keys) {
  const unsigned long long key = keys[blockIdx.y];
  double* o = out + (int64_t)blockIdx.y * per;
  const int64_t step = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < per; i += step) {
    unsigned long long z = key + (unsigned long long)i * 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    z = z ^ (z >> 31);
    o[i] = (double)(z >> 11) * (1.0 / 9007199254740992.0);
  }
}

// kept-rank rule for matrix blockIdx.x (truncation_count_kernel with strides)
__global__ void truncation_count_batched_kernel(const double* s, int64_t n, int64_t s_stride, int64_t chi, double threshold,
                                                int relative, double* info, double* s_scaled) {
  const int64_t b = blockIdx.x;
  s += b * s_stride;
  if (s_scaled) s_scaled += b * s_stride;
  __shared__ int cnt;
  if (threadIdx.x == 0) cnt = 0;
  __syncthreads();
  const double s0 = n > 0 ? s[0] : 0.0;
  const double bar = relative ? threshold * s0 : threshold;
  const int64_t lim = (chi > 0 && chi < n) ? chi : n;
  int local = 0;
  for (int64_t i = threadIdx.x; i < n; i += blockDim.x) {
    const double v = s[i];
    const bool above = (relative == 2) ? (v / s0 > threshold) : (v > bar);
    if (i < lim && above) local++;
    if (s_scaled) s_scaled[i] = v / s0;
  }
  atomicAdd(&cnt, local);
  __syncthreads();
  if (threadIdx.x == 0) { info[2 * b] = (double)cnt; info[2 * b + 1] = s0; }
}

// out[b] = |x[b]|_F of `per` contiguous elements, one CTA per matrix (amax-scaled: no overflow)
template <typename T>
__global__ void __launch_bounds__(512) norm2_batched_kernel(const T* x, int64_t per, int64_t stride, double* out) {
  const T* xb = x + (int64_t)blockIdx.x * stride;
  __shared__ double red[16];
  __shared__ double bc;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  double mx = 0.0;
  for (int64_t i = tid; i < per; i += 512) {
    const double a = sqrt(Num<T>::abs2(xb[i]));
    mx = (a != a) ? a : fmax(mx, a);
    if (a != a) break;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const double other = __shfl_xor_sync(0xffffffffu, mx, o);
    mx = (mx != mx) ? mx : ((other != other) ? other : fmax(mx, other));
  }
  if (lane == 0) red[warp] = mx;
  __syncthreads();
  if (tid == 0) {
    double m_ = 0.0;
    for (int w = 0; w < 16; ++w) m_ = (m_ != m_) ? m_ : ((red[w] != red[w]) ? red[w] : fmax(m_, red[w]));
    bc = m_;
  }
  __syncthreads();
  const double amax = bc;
  if (!(amax > 0.0) || !isfinite(amax)) { if (tid == 0) out[blockIdx.x] = amax; return; }
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
    out[blockIdx.x] = amax * sqrt(t);
  }
}
}  // namespace tnb

extern "C" int tnb_fill_uniform_batched(double* out, int64_t per, int64_t batch, const unsigned long long* keys_device,
                                        void* stream) {
  if (!out || !keys_device || per < 0 || batch < 0 || batch > 65535) return TNB_E_ARG;
  if (per == 0 || batch == 0) return 0;
  unsigned gx = grid_for(per);
  if (gx > 256) gx = 256;
  tnb::fill_uniform_batched_kernel<<<dim3(gx, (unsigned)batch), 256, 0, (cudaStream_t)stream>>>(out, per, keys_device);
  TNB_LAUNCH_CHECK();
  return 0;
}

extern "C" int tnb_truncation_count_batched(const double* s, int64_t n, int64_t stride, int64_t batch, int64_t chi,
                                            double threshold, int relative, double* info_device, double* s_scaled,
                                            void* stream) {
  if (!s || !info_device || n < 0 || batch < 0 || stride < n) return TNB_E_ARG;
  if (batch == 0) return 0;
  tnb::truncation_count_batched_kernel<<<(unsigned)batch, 256, 0, (cudaStream_t)stream>>>(s, n, stride, chi, threshold,
                                                                                        relative, info_device, s_scaled);
  TNB_LAUNCH_CHECK();
  return 0;
}

extern "C" int tnb_norm2_batched(int dtype, const void* x, int64_t per, int64_t stride, int64_t batch, double* out_device,
                                 void* stream) {
  if (!x || !out_device || per < 0 || batch < 0 || stride < per) return TNB_E_ARG;
  if (batch == 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  if (per == 0) { TNB_CUDA_CHECK(cudaMemsetAsync(out_device, 0, sizeof(double) * (size_t)batch, st)); return 0; }
  ProfScope prof(KC_ELEMWISE, st, (double)per * (double)batch * (double)elem_size(dtype));
  if (dtype == TNB_F64) tnb::norm2_batched_kernel<double><<<(unsigned)batch, 512, 0, st>>>((const double*)x, per, stride, out_device);
  else if (dtype == TNB_C128) tnb::norm2_batched_kernel<double2><<<(unsigned)batch, 512, 0, st>>>((const double2*)x, per, stride, out_device);
  else return TNB_E_ARG;
  TNB_LAUNCH_CHECK();
  return 0;
}
