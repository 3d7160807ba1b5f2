// Index permutation (label -> axis reordering) for float64 / complex128.
//
// out (contiguous) = alpha * op(in viewed through `perm`).  HBM-bound:
// algorithmic traffic = 2 * numel * sizeof(element).  Two kernels:
//   * permute_rows   -- the input's unit-stride axis is also the output's last
//                       axis (or nothing has unit stride): every thread moves
//                       one 16-byte packet, reads and writes both coalesced
//                       along the inner run;
//   * permute_tiled  -- the input's unit-stride axis lands somewhere else in
//                       the output: 32x32 tiles staged through padded shared
//                       memory so both the global read (along the input-fast
//                       axis) and the global write (along the output-fast axis)
//                       are coalesced.
// Replaces the hidden copies of np.rollaxis + np.reshape in the reference
// (tensor.py:295,315,354,363,392-394,482,818,911,1041).
#include "common.cuh"

namespace tnb {

std::atomic<long long> g_launches{0};
thread_local int g_pdl = 0;
thread_local const int* g_skip_flag = nullptr;
// SM count of the CURRENT device, cached per device (a process may drive several GPUs, from several threads)
int sm_count() {
  static std::atomic<int> cache[64];
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 0 || dev >= 64) dev = 0;
  int n = cache[dev].load(std::memory_order_relaxed);
  if (n == 0) {
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
    cache[dev].store(n, std::memory_order_relaxed);
  }
  return n;
}

struct PermParams {
  int rank;
  int a_axis;  // tiled kernel: output axis that is unit-stride in the input
  int64_t oshape[TNB_MAX_RANK];
  int64_t istride[TNB_MAX_RANK];
  int64_t ostride[TNB_MAX_RANK];
  int64_t numel;
  double ar, ai;
  int conj;
  int scale;  // 0: plain copy, 1: multiply by alpha
};

// MODE 0: double, 1: complex128, 2: pair of doubles moved as one 16-byte packet
template <int MODE> struct Pack;
template <> struct Pack<0> { typedef double type; };
template <> struct Pack<1> { typedef double2 type; };
template <> struct Pack<2> { typedef double2 type; };

template <int MODE>
__device__ __forceinline__ typename Pack<MODE>::type apply(typename Pack<MODE>::type v, const PermParams& p) {
  if constexpr (MODE == 0) {
    return p.scale ? v * p.ar : v;
  } else if constexpr (MODE == 1) {
    if (p.conj) v.y = -v.y;
    if (p.scale) v = cmul(v, make_double2(p.ar, p.ai));
    return v;
  } else {
    if (p.scale) { v.x *= p.ar; v.y *= p.ar; }
    return v;
  }
}

// Every thread moves PR_ILP packets per iteration (a CTA covers PR_ILP consecutive runs of 256 packets): the
// index decomposition of the four packets is interleaved and their loads are all in flight before the first
// store -- the kernel is latency-bound otherwise (one dependent chain of divisions + one load per thread).
constexpr int PR_ILP = 4;
template <int MODE, typename IDX>
__global__ void __launch_bounds__(256) permute_rows(const typename Pack<MODE>::type* __restrict__ in,
                                                    typename Pack<MODE>::type* __restrict__ out, PermParams p) {
  typedef typename Pack<MODE>::type T;
  const IDX n = (IDX)p.numel;
  const IDX step = (IDX)gridDim.x * blockDim.x * PR_ILP;
  for (IDX base = (IDX)blockIdx.x * blockDim.x * PR_ILP + threadIdx.x; base < n; base += step) {
    IDX rem[PR_ILP];
    int64_t off[PR_ILP];
#pragma unroll
    for (int u = 0; u < PR_ILP; ++u) {
      const IDX idx = base + (IDX)u * blockDim.x;
      rem[u] = idx < n ? idx : (IDX)0;
      off[u] = 0;
    }
#pragma unroll 1
    for (int d = p.rank - 1; d > 0; --d) {
      const IDX s = (IDX)p.oshape[d];
      const int64_t is = p.istride[d];
#pragma unroll
      for (int u = 0; u < PR_ILP; ++u) {
        const IDX q = rem[u] / s;
        off[u] += (int64_t)(rem[u] - q * s) * is;
        rem[u] = q;
      }
    }
    T v[PR_ILP];
#pragma unroll
    for (int u = 0; u < PR_ILP; ++u) v[u] = in[off[u] + (int64_t)rem[u] * p.istride[0]];
#pragma unroll
    for (int u = 0; u < PR_ILP; ++u) {
      const IDX idx = base + (IDX)u * blockDim.x;
      if (idx < n) out[idx] = apply<MODE>(v[u], p);
    }
  }
}

// 32x32 tile transpose with batch dims.  block = (32, 8).
template <int MODE>
__global__ void __launch_bounds__(256) permute_tiled(const typename Pack<MODE>::type* __restrict__ in,
                                                     typename Pack<MODE>::type* __restrict__ out, PermParams p,
                                                     int64_t tiles_a, int64_t tiles_b) {
  typedef typename Pack<MODE>::type T;
  __shared__ T tile[32][33];
  const int last = p.rank - 1;
  const int a = p.a_axis;
  int64_t t = blockIdx.x;
  const int64_t tb = t % tiles_b; t /= tiles_b;
  const int64_t ta = t % tiles_a; t /= tiles_a;
  // t now indexes the remaining axes (all but a and last), row-major
  int64_t ibase = 0, obase = 0;
  for (int d = last - 1; d >= 0; --d) {
    if (d == a) continue;
    const int64_t s = p.oshape[d];
    const int64_t r = t % s;
    t /= s;
    ibase += r * p.istride[d];
    obase += r * p.ostride[d];
  }
  const int64_t Na = p.oshape[a], Nb = p.oshape[last];
  const int64_t sb_in = p.istride[last], sa_out = p.ostride[a];
  const int tx = threadIdx.x, ty = threadIdx.y;
  {
    const int64_t ai = ta * 32 + tx;
#pragma unroll
    for (int j = ty; j < 32; j += 8) {
      const int64_t bi = tb * 32 + j;
      if (ai < Na && bi < Nb) tile[j][tx] = in[ibase + ai + bi * sb_in];
    }
  }
  __syncthreads();
  {
    const int64_t bi = tb * 32 + tx;
#pragma unroll
    for (int j = ty; j < 32; j += 8) {
      const int64_t ai = ta * 32 + j;
      if (ai < Na && bi < Nb) out[obase + ai * sa_out + bi] = apply<MODE>(tile[tx][j], p);
    }
  }
}

template <int MODE>
static int launch_permute(const void* in, void* out, PermParams& p, cudaStream_t st) {
  typedef typename Pack<MODE>::type T;
  if (p.numel == 0) return 0;
  const int last = p.rank - 1;
  // find the output axis that is unit-stride in the input
  int a = -1;
  for (int d = 0; d < p.rank; ++d)
    if (p.istride[d] == 1 && p.oshape[d] > 1) a = d;
  if (a >= 0 && a != last && p.oshape[a] >= 8 && p.oshape[last] >= 8) {
    p.a_axis = a;
    const int64_t tiles_a = (p.oshape[a] + 31) / 32, tiles_b = (p.oshape[last] + 31) / 32;
    int64_t rest = 1;
    for (int d = 0; d < last; ++d)
      if (d != a) rest *= p.oshape[d];
    const int64_t nblk = tiles_a * tiles_b * rest;
    if (nblk < (int64_t)2147483647) {
      permute_tiled<MODE><<<(unsigned)nblk, dim3(32, 8), 0, st>>>((const T*)in, (T*)out, p, tiles_a, tiles_b);
      TNB_LAUNCH_CHECK();
      return 0;
    }
  }
  const int threads = 256;
  int64_t blocks = (p.numel + threads * PR_ILP - 1) / (threads * PR_ILP);
  const int64_t cap = (int64_t)sm_count() * 32;
  if (blocks > cap) blocks = cap;
  if (p.numel < (int64_t)2000000000)
    permute_rows<MODE, uint32_t><<<(unsigned)blocks, threads, 0, st>>>((const T*)in, (T*)out, p);
  else
    permute_rows<MODE, int64_t><<<(unsigned)blocks, threads, 0, st>>>((const T*)in, (T*)out, p);
  TNB_LAUNCH_CHECK();
  return 0;
}

// Shared by tensordot.cu / qr.cu / svd.cu: out = alpha*op(view), view given in OUTPUT axis order.
int permute_view(int dtype, const void* in, int rank, const int64_t* oshape, const int64_t* istride, void* out,
                 double ar, double ai, int conj, cudaStream_t st) {
  View v = collapse(rank, oshape, istride);
  PermParams p;
  p.rank = v.rank;
  p.numel = v.numel;
  p.a_axis = -1;
  for (int d = 0; d < v.rank; ++d) { p.oshape[d] = v.shape[d]; p.istride[d] = v.stride[d]; }
  int64_t s = 1;
  for (int d = v.rank - 1; d >= 0; --d) { p.ostride[d] = s; s *= p.oshape[d]; }
  p.ar = ar; p.ai = ai;
  p.conj = (dtype == TNB_C128) ? (conj != 0) : 0;
  p.scale = !(ar == 1.0 && ai == 0.0);
  ProfScope prof(KC_PERMUTE, st, 2.0 * (double)v.numel * (double)elem_size(dtype));
  if (dtype == TNB_C128) return launch_permute<1>(in, out, p, st);
  // float64: move pairs as 16-byte packets when the inner run allows it
  const int last = p.rank - 1;
  bool pair = p.istride[last] == 1 && (p.oshape[last] % 2 == 0) && ((uintptr_t)in % 16 == 0) &&
              ((uintptr_t)out % 16 == 0);
  for (int d = 0; d < last && pair; ++d) pair = (p.istride[d] % 2 == 0);
  if (pair) {
    PermParams q = p;
    q.oshape[last] /= 2;
    for (int d = 0; d < last; ++d) { q.istride[d] /= 2; q.ostride[d] /= 2; }
    q.numel /= 2;
    return launch_permute<2>(in, out, q, st);
  }
  return launch_permute<0>(in, out, p, st);
}

}  // namespace tnb

extern "C" int tnb_permute(const tnb_tensor_t* in, const int32_t* perm, void* out, double alpha_re, double alpha_im,
                           int conj, void* stream) {
  using namespace tnb;
  if (!valid_tensor(in) || !out || (in->rank > 0 && !perm)) return TNB_E_ARG;
  int64_t oshape[TNB_MAX_RANK], istride[TNB_MAX_RANK];
  bool seen[TNB_MAX_RANK] = {false};
  for (int i = 0; i < in->rank; ++i) {
    const int a = perm[i];
    if (a < 0 || a >= in->rank || seen[a]) return TNB_E_ARG;
    seen[a] = true;
    oshape[i] = in->shape[a];
    istride[i] = in->stride[a];
  }
  if (in->dtype == TNB_F64 && alpha_im != 0.0) return TNB_E_ARG;
  return permute_view(in->dtype, in->ptr, in->rank, oshape, istride, out, alpha_re, alpha_im, conj,
                      (cudaStream_t)stream);
}
