// Shared device/host helpers for libtnb (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include <atomic>

#include "../../include/tnb.h"

#define TNB_CUDA_CHECK(expr)                         \
  do {                                               \
    cudaError_t _e = (expr);                         \
    if (_e != cudaSuccess) return (int)_e;           \
  } while (0)

// every kernel launch of the library goes through this macro, so the counter
// is the number of libtnb kernels launched (bench.py reports it as gpu_launches)
namespace tnb {
extern std::atomic<long long> g_launches;   // several host threads may drive the library (one stream each)
inline void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }
}
#define TNB_LAUNCH_CHECK()                   \
  do {                                       \
    tnb::count_launch();                     \
    TNB_CUDA_CHECK(cudaGetLastError());      \
  } while (0)

namespace tnb {

// kernel classes of the optional device-time profile (prof.cu, tnb_profile_*)
enum { KC_GEMM = 0, KC_JACOBI = 1, KC_QR_PANEL = 2, KC_PERMUTE = 3, KC_MPO_APPLY = 4, KC_ELEMWISE = 5, KC_COUNT = 6 };
extern bool g_prof_on;
// RAII: brackets the kernel launches made during its lifetime with two CUDA
// events on `st` when profiling is on; `work` = algorithmic flops or bytes.
struct ProfScope {
  int cls;
  cudaStream_t st;
  double work;
  long long l0, idx;
  ProfScope(int c, cudaStream_t s, double w);
  ~ProfScope();
};

typedef double2 cplx;  // interleaved complex128

__host__ __device__ __forceinline__ cplx cmul(cplx a, cplx b) {
  return make_double2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}
__host__ __device__ __forceinline__ cplx cadd(cplx a, cplx b) { return make_double2(a.x + b.x, a.y + b.y); }
__host__ __device__ __forceinline__ cplx csub(cplx a, cplx b) { return make_double2(a.x - b.x, a.y - b.y); }
__host__ __device__ __forceinline__ cplx cconj(cplx a) { return make_double2(a.x, -a.y); }
__host__ __device__ __forceinline__ cplx cscale(cplx a, double s) { return make_double2(a.x * s, a.y * s); }
// a + conj(x)*y
__host__ __device__ __forceinline__ cplx cfma_conj(cplx x, cplx y, cplx a) {
  return make_double2(a.x + x.x * y.x + x.y * y.y, a.y + x.x * y.y - x.y * y.x);
}
// a + x*y
__host__ __device__ __forceinline__ cplx cfma(cplx x, cplx y, cplx a) {
  return make_double2(a.x + x.x * y.x - x.y * y.y, a.y + x.x * y.y + x.y * y.x);
}
__host__ __device__ __forceinline__ double cabs2(cplx a) { return a.x * a.x + a.y * a.y; }

// Scalar traits so kernels can be written once for double and complex128.
template <typename T> struct Num;
template <> struct Num<double> {
  typedef double real_t;
  static constexpr int dtype = TNB_F64;
  __host__ __device__ static double zero() { return 0.0; }
  __host__ __device__ static double one() { return 1.0; }
  __host__ __device__ static double conj(double a) { return a; }
  __host__ __device__ static double mul(double a, double b) { return a * b; }
  __host__ __device__ static double add(double a, double b) { return a + b; }
  __host__ __device__ static double sub(double a, double b) { return a - b; }
  __host__ __device__ static double scale(double a, double s) { return a * s; }
  __host__ __device__ static double abs2(double a) { return a * a; }
  __host__ __device__ static double real(double a) { return a; }
  __host__ __device__ static double from(double re, double) { return re; }
  __host__ __device__ static double fma_conj(double x, double y, double a) { return a + x * y; }
  __host__ __device__ static double fma(double x, double y, double a) { return a + x * y; }
};
template <> struct Num<cplx> {
  typedef double real_t;
  static constexpr int dtype = TNB_C128;
  __host__ __device__ static cplx zero() { return make_double2(0.0, 0.0); }
  __host__ __device__ static cplx one() { return make_double2(1.0, 0.0); }
  __host__ __device__ static cplx conj(cplx a) { return cconj(a); }
  __host__ __device__ static cplx mul(cplx a, cplx b) { return cmul(a, b); }
  __host__ __device__ static cplx add(cplx a, cplx b) { return cadd(a, b); }
  __host__ __device__ static cplx sub(cplx a, cplx b) { return csub(a, b); }
  __host__ __device__ static cplx scale(cplx a, double s) { return cscale(a, s); }
  __host__ __device__ static double abs2(cplx a) { return cabs2(a); }
  __host__ __device__ static double real(cplx a) { return a.x; }
  __host__ __device__ static cplx from(double re, double im) { return make_double2(re, im); }
  __host__ __device__ static cplx fma_conj(cplx x, cplx y, cplx a) { return cfma_conj(x, y, a); }
  __host__ __device__ static cplx fma(cplx x, cplx y, cplx a) { return cfma(x, y, a); }
};

// Single-precision counterparts: used ONLY for the Gram-domain rotation phase of the Jacobi round (svd.cu), where the
// data decides angles but never touches the matrix itself.
template <> struct Num<float> {
  typedef float real_t;
  __host__ __device__ static float zero() { return 0.f; }
  __host__ __device__ static float one() { return 1.f; }
  __host__ __device__ static float conj(float a) { return a; }
  __host__ __device__ static float add(float a, float b) { return a + b; }
  __host__ __device__ static float sub(float a, float b) { return a - b; }
  __host__ __device__ static float scale(float a, float s) { return a * s; }
  __host__ __device__ static float abs2(float a) { return a * a; }
  __host__ __device__ static float real(float a) { return a; }
  __host__ __device__ static float from(float re, float) { return re; }
};
template <> struct Num<float2> {
  typedef float real_t;
  __host__ __device__ static float2 zero() { return make_float2(0.f, 0.f); }
  __host__ __device__ static float2 one() { return make_float2(1.f, 0.f); }
  __host__ __device__ static float2 conj(float2 a) { return make_float2(a.x, -a.y); }
  __host__ __device__ static float2 add(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
  __host__ __device__ static float2 sub(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }
  __host__ __device__ static float2 scale(float2 a, float s) { return make_float2(a.x * s, a.y * s); }
  __host__ __device__ static float abs2(float2 a) { return fmaf(a.x, a.x, a.y * a.y); }
  __host__ __device__ static float real(float2 a) { return a.x; }
  __host__ __device__ static float2 from(float re, float im) { return make_float2(re, im); }
};
template <typename T> struct LowPrec;
template <> struct LowPrec<double> { typedef float type; };
template <> struct LowPrec<cplx> { typedef float2 type; };

static inline size_t elem_size(int dtype) { return dtype == TNB_C128 ? 16 : 8; }

// ---- PTX wrappers -----------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}
// cp.async with zero-fill: copies src_bytes (<= CP) and zero-fills the rest.
__device__ __forceinline__ void cp_async16(void* smem, const void* gmem, int src_bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(smem_u32(smem)), "l"(gmem), "r"(src_bytes));
}
__device__ __forceinline__ void cp_async8(void* smem, const void* gmem, int src_bytes) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;\n" ::"r"(smem_u32(smem)), "l"(gmem), "r"(src_bytes));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N> __device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;\n" ::"n"(N));
}
// D(8x8) += A(8x4, row) * B(4x8, col): one DMMA.8x8x4 on sm_100a.
__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(c0), "+d"(c1)
               : "d"(a), "d"(b));
}

// ---- TMA 1-D bulk copies (cp.async.bulk, SASS UBLKCP) + mbarrier completion -------------------------
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory"); }
// generic-proxy shared-memory writes -> visible to the async proxy (bulk stores read smem through it)
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}\n"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  while (!mbar_try_wait(bar, parity)) {}
}
// global -> shared, `bytes` a multiple of 16, both addresses 16-byte aligned; completes on `bar`
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(
                   smem_u32(smem_dst)),
               "l"(gsrc), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
// shared -> global (bulk async-group completion)
__device__ __forceinline__ void bulk_s2g(void* gdst, const void* smem_src, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;\n" ::"l"(gdst), "r"(smem_u32(smem_src)),
               "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;\n" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;\n" ::: "memory"); }
// waits only until the shared-memory SOURCE of every committed bulk store has been read (the buffer may
// be reused, the CTA may exit); the global writes themselves complete by the end of the grid at the latest
__device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;\n" ::: "memory"); }

// ---- programmatic dependent launch (launch attribute cudaLaunchAttributeProgrammaticStreamSerialization) ----
// wait: the preceding grid of the stream has completed and its writes are visible; launch_dependents: the
// next grid may be scheduled once every CTA of this one has executed it (or exited).
__device__ __forceinline__ void griddep_wait() { asm volatile("griddepcontrol.wait;\n" ::: "memory"); }
__device__ __forceinline__ void griddep_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;\n" ::: "memory"); }

// Host side of programmatic dependent launch for kernel chains: inside a PdlScope every launch made through
// launch_k() carries the programmatic-stream-serialization attribute.  Only kernels that execute griddep_wait()
// before their first global-memory access may be launched this way (the GEMM, split-K reduce, Cholesky and copy
// kernels do).  Measured on the Gram-Schmidt QR chain (3072 x 1536 complex128): 7.40 ms with the scope against
// 6.71 without -- the early-scheduled CTAs of the next short kernel take SM slots from the side-stream GEMM -- so
// no caller opens a scope today; the Jacobi rounds carry the attribute themselves (svd.cu launch_round).
extern thread_local int g_pdl;
struct PdlScope {
  bool on;
  explicit PdlScope(bool enable) : on(enable) { if (on) ++g_pdl; }
  ~PdlScope() { if (on) --g_pdl; }
};
// Device-side skip: kernels that take the flag (GEMM, split-K reduce, copy2d) return at once when *flag != 0.  The
// lagged second pass of the Gram-Schmidt QR queues its update kernels inside a SkipScope and lets a device
// predicate decide whether they run (no host round trip).
extern thread_local const int* g_skip_flag;
struct SkipScope {
  const int* prev;
  explicit SkipScope(const int* f) : prev(g_skip_flag) { g_skip_flag = f; }
  ~SkipScope() { g_skip_flag = prev; }
};
template <typename... KArgs, typename... Args>
static inline cudaError_t launch_k(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = g_pdl > 0 ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kern, KArgs(args)...);
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// ---- collapsed strided view used by permute / elementwise kernels ------------
struct View {
  int rank;
  int64_t shape[TNB_MAX_RANK];
  int64_t stride[TNB_MAX_RANK];
  int64_t numel;
};

// Drop size-1 axes and merge axes that are contiguous with respect to each
// other (stride[i] == stride[i+1]*shape[i+1]); logical (row-major) order kept.
static inline View collapse(int rank, const int64_t* shape, const int64_t* stride) {
  View v;
  v.rank = 0;
  v.numel = 1;
  for (int i = 0; i < rank; ++i) {
    v.numel *= shape[i];
    if (shape[i] == 1) continue;
    if (v.rank > 0 && v.stride[v.rank - 1] == stride[i] * shape[i]) {
      v.shape[v.rank - 1] *= shape[i];
      v.stride[v.rank - 1] = stride[i];
    } else {
      v.shape[v.rank] = shape[i];
      v.stride[v.rank] = stride[i];
      v.rank++;
    }
  }
  if (v.rank == 0) {
    v.rank = 1;
    v.shape[0] = 1;
    v.stride[0] = 1;
  }
  return v;
}

static inline bool valid_tensor(const tnb_tensor_t* t) {
  if (!t || t->rank < 0 || t->rank > TNB_MAX_RANK) return false;
  if (t->dtype != TNB_F64 && t->dtype != TNB_C128) return false;
  for (int i = 0; i < t->rank; ++i)
    if (t->shape[i] < 0) return false;
  return true;
}
static inline int64_t numel(const tnb_tensor_t* t) {
  int64_t n = 1;
  for (int i = 0; i < t->rank; ++i) n *= t->shape[i];
  return n;
}

int sm_count();  // cached cudaDevAttrMultiProcessorCount of the current device

// "This per-device function attribute has been set on the current device": one mark per device, safe against
// concurrent first use from several host threads (setting an attribute twice is harmless, so a plain atomic
// flag is enough).  Usage:  static PerDeviceOnce once;  if (once.need()) { cudaFuncSetAttribute(...); once.done(); }
struct PerDeviceOnce {
  std::atomic<int> mark[64];
  static int dev() { int d = 0; cudaGetDevice(&d); return (d >= 0 && d < 64) ? d : 0; }
  bool need() const { return mark[dev()].load(std::memory_order_acquire) == 0; }
  void done() { mark[dev()].store(1, std::memory_order_release); }
};

}  // namespace tnb
