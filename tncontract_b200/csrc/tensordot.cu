// np.tensordot on device (reference call site: tensor.py:735).
//
// NumPy transposes both operands to (free|contracted) / (contracted|free),
// materialises the copies and calls BLAS gemm.  Here the planner first tries
// to express each operand AS IT LIES IN MEMORY as one of the GEMM operand
// forms the kernel can stage directly:
//     K-contiguous   [mn][k]   (ld = stride of the merged free group)
//     MN-contiguous  [k][mn]   (ld = stride of the merged contracted group)
//     either of the above with ONE leading free axis peeled off as a batch
//     index (e.g. contracting the middle axis of A[phys,left,right])
// and only falls back to a permutation copy (permute.cu) into the caller's
// workspace when no such form exists.  All sweep-path contractions of the MPS
// routines hit a copy-free form.
#include "common.cuh"

namespace tnb {

int permute_view(int dtype, const void* in, int rank, const int64_t* oshape, const int64_t* istride, void* out,
                 double ar, double ai, int conj, cudaStream_t st);
int gemm_ws(int dtype, int opA, int opB, int64_t M, int64_t N, int64_t K, double ar, double ai, const void* A,
            int64_t lda, int64_t sA, const void* B, int64_t ldb, int64_t sB, double br, double bi, void* C, int64_t ldc,
            int64_t sC, int64_t batch, void* splitk_ws, size_t splitk_bytes, cudaStream_t st);
int gemm(int dtype, int opA, int opB, int64_t M, int64_t N, int64_t K, double ar, double ai, const void* A,
         int64_t lda, int64_t sA, const void* B, int64_t ldb, int64_t sB, double br, double bi, void* C, int64_t ldc,
         int64_t sC, int64_t batch, cudaStream_t st);

struct Group {
  bool ok;
  int64_t size, stride;  // stride meaningless when size == 1
};

static Group merge_axes(const tnb_tensor_t* t, const int* axes, int n) {
  Group g = {true, 1, 0};
  bool have = false;
  for (int i = 0; i < n; ++i) {
    const int64_t s = t->shape[axes[i]], st = t->stride[axes[i]];
    if (s == 1) continue;
    if (!have) { g.size = s; g.stride = st; have = true; continue; }
    if (g.stride != st * s) { g.ok = false; return g; }
    g.size *= s;
    g.stride = st;
  }
  return g;
}

struct OperandPlan {
  bool ok = false;
  bool kc = false;      // true: memory is [mn][k] (k contiguous); false: [k][mn]
  int64_t ld = 1;
  int64_t mn = 1;       // MN extent seen by one GEMM of the batch
  int64_t nbatch = 1, bstride = 0;
};

static bool plan_2d(const Group& mn, const Group& k, OperandPlan* p) {
  if (!mn.ok || !k.ok) return false;
  p->mn = mn.size;
  if (k.size == 1 || k.stride == 1) {
    p->kc = true;
    p->ld = (mn.size == 1) ? (k.size > 0 ? k.size : 1) : mn.stride;
    return true;
  }
  if (mn.size == 1 || mn.stride == 1) {
    p->kc = false;
    p->ld = k.stride;
    return true;
  }
  return false;
}

static OperandPlan analyze(const tnb_tensor_t* t, const int* mn_axes, int n_mn, const int* k_axes, int n_k,
                           bool allow_batch) {
  OperandPlan p;
  const Group k = merge_axes(t, k_axes, n_k);
  if (!k.ok) return p;
  const Group mn = merge_axes(t, mn_axes, n_mn);
  if (plan_2d(mn, k, &p)) { p.ok = true; return p; }
  if (!allow_batch) return p;
  // peel the first non-unit free axis as the batch index
  int first = -1;
  for (int i = 0; i < n_mn; ++i)
    if (t->shape[mn_axes[i]] > 1) { first = i; break; }
  if (first < 0) return p;
  const Group rest = merge_axes(t, mn_axes + first + 1, n_mn - first - 1);
  OperandPlan q;
  if (plan_2d(rest, k, &q)) {
    q.ok = true;
    q.nbatch = t->shape[mn_axes[first]];
    q.bstride = t->stride[mn_axes[first]];
    return q;
  }
  return p;
}

struct DotPlan {
  int64_t M, N, K;
  int nfa, nfb;
  int fa[TNB_MAX_RANK], fb[TNB_MAX_RANK], ca[TNB_MAX_RANK], cb[TNB_MAX_RANK];
  OperandPlan pa, pb;
  bool copy_a, copy_b;
  size_t ws_a, ws_b;
  size_t ws_sk;   // split-K scratch (complex products between one and two waves of full-size tiles, see gemm_ws)
};

static int make_plan(const tnb_tensor_t* a, const tnb_tensor_t* b, int nctr, const int32_t* axes_a,
                     const int32_t* axes_b, DotPlan* P) {
  if (!valid_tensor(a) || !valid_tensor(b) || nctr < 0 || nctr > TNB_MAX_RANK) return TNB_E_ARG;
  if (a->dtype != b->dtype) return TNB_E_ARG;
  if (nctr > 0 && (!axes_a || !axes_b)) return TNB_E_ARG;
  bool ua[TNB_MAX_RANK] = {false}, ub[TNB_MAX_RANK] = {false};
  P->K = 1;
  for (int i = 0; i < nctr; ++i) {
    const int x = axes_a[i], y = axes_b[i];
    if (x < 0 || x >= a->rank || y < 0 || y >= b->rank || ua[x] || ub[y]) return TNB_E_ARG;
    if (a->shape[x] != b->shape[y]) return TNB_E_ARG;
    ua[x] = ub[y] = true;
    P->K *= a->shape[x];
  }
  P->nfa = P->nfb = 0;
  P->M = P->N = 1;
  for (int i = 0; i < a->rank; ++i)
    if (!ua[i]) { P->fa[P->nfa++] = i; P->M *= a->shape[i]; }
  for (int i = 0; i < b->rank; ++i)
    if (!ub[i]) { P->fb[P->nfb++] = i; P->N *= b->shape[i]; }
  if (a->rank + b->rank - 2 * nctr > TNB_MAX_RANK) return TNB_E_UNSUPPORTED;

  // Candidate orders of the contraction pairs: as given, by A's stride, by
  // B's stride (descending).  Any common reordering of the pairs is valid.
  int best_cost = 100;
  for (int cand = 0; cand < 3; ++cand) {
    int order[TNB_MAX_RANK];
    for (int i = 0; i < nctr; ++i) order[i] = i;
    if (cand > 0) {
      const tnb_tensor_t* t = (cand == 1) ? a : b;
      const int32_t* ax = (cand == 1) ? axes_a : axes_b;
      for (int i = 1; i < nctr; ++i)  // insertion sort, stride descending
        for (int j = i; j > 0 && t->stride[ax[order[j]]] > t->stride[ax[order[j - 1]]]; --j) {
          const int tmp = order[j]; order[j] = order[j - 1]; order[j - 1] = tmp;
        }
    }
    int ca[TNB_MAX_RANK], cb[TNB_MAX_RANK];
    for (int i = 0; i < nctr; ++i) { ca[i] = axes_a[order[i]]; cb[i] = axes_b[order[i]]; }
    OperandPlan pa = analyze(a, P->fa, P->nfa, ca, nctr, true);
    OperandPlan pb = analyze(b, P->fb, P->nfb, cb, nctr, true);
    // only one operand may carry the batch index
    if (pa.ok && pb.ok && pa.nbatch > 1 && pb.nbatch > 1) {
      OperandPlan pa2 = analyze(a, P->fa, P->nfa, ca, nctr, false);
      OperandPlan pb2 = analyze(b, P->fb, P->nfb, cb, nctr, false);
      if (pb2.ok) pb = pb2;
      else if (pa2.ok) pa = pa2;
      else if (numel(a) <= numel(b)) pa = pa2;
      else pb = pb2;
    }
    const int cost = (pa.ok ? 0 : 1) + (pb.ok ? 0 : 1);
    if (cost < best_cost) {
      best_cost = cost;
      P->pa = pa; P->pb = pb;
      for (int i = 0; i < nctr; ++i) { P->ca[i] = ca[i]; P->cb[i] = cb[i]; }
    }
    if (best_cost == 0) break;
  }
  P->copy_a = !P->pa.ok;
  P->copy_b = !P->pb.ok;
  // a batched operand needs the OTHER operand un-batched (always true after
  // the loop above), and a copied operand is never batched.
  const size_t es = elem_size(a->dtype);
  P->ws_a = P->copy_a ? (((size_t)numel(a) * es + 255) & ~(size_t)255) : 0;
  P->ws_b = P->copy_b ? (((size_t)numel(b) * es + 255) & ~(size_t)255) : 0;
  // un-batched complex products whose full-size tiles fill between one and two waves get room for four partials
  P->ws_sk = 0;
  const bool batched = (P->pa.ok && P->pa.nbatch > 1) || (P->pb.ok && P->pb.nbatch > 1);
  if (a->dtype == TNB_C128 && !batched && P->K >= 512) {
    const int64_t tiles = ((P->M + 127) / 128) * ((P->N + 63) / 64), cap = sm_count();
    if (tiles > cap && tiles < 2 * cap) P->ws_sk = (4 * (size_t)P->M * P->N * es + 255) & ~(size_t)255;
  }
  return 0;
}

}  // namespace tnb

using namespace tnb;

extern "C" size_t tnb_tensordot_workspace(const tnb_tensor_t* a, const tnb_tensor_t* b, int nctr,
                                          const int32_t* axes_a, const int32_t* axes_b) {
  DotPlan P;
  if (make_plan(a, b, nctr, axes_a, axes_b, &P) != 0) return 0;
  return P.ws_a + P.ws_b + P.ws_sk;
}

extern "C" int tnb_tensordot(const tnb_tensor_t* a, const tnb_tensor_t* b, int nctr, const int32_t* axes_a,
                             const int32_t* axes_b, int conj_a, int conj_b, void* out, void* ws, size_t ws_bytes,
                             void* stream) {
  DotPlan P;
  int rc = make_plan(a, b, nctr, axes_a, axes_b, &P);
  if (rc != 0) return rc;
  if (P.M == 0 || P.N == 0) return 0;
  if (!out) return TNB_E_ARG;
  if (P.ws_a + P.ws_b > ws_bytes || (P.ws_a + P.ws_b > 0 && !ws)) return TNB_E_WORKSPACE;
  const size_t sk_bytes = (ws && ws_bytes >= P.ws_a + P.ws_b + P.ws_sk) ? P.ws_sk : 0;   // optional: a caller may pass less
  void* const sk = sk_bytes ? (char*)ws + P.ws_a + P.ws_b : nullptr;
  cudaStream_t st = (cudaStream_t)stream;
  const int dtype = a->dtype;
  const bool cplx = dtype == TNB_C128;
  const void* Aptr = a->ptr;
  const void* Bptr = b->ptr;
  OperandPlan pa = P.pa, pb = P.pb;
  if (P.copy_a) {  // -> [free..., ctr...] contiguous: K-contiguous, ld = K
    int64_t sh[TNB_MAX_RANK], is[TNB_MAX_RANK];
    int r = 0;
    for (int i = 0; i < P.nfa; ++i, ++r) { sh[r] = a->shape[P.fa[i]]; is[r] = a->stride[P.fa[i]]; }
    for (int i = 0; i < nctr; ++i, ++r) { sh[r] = a->shape[P.ca[i]]; is[r] = a->stride[P.ca[i]]; }
    rc = permute_view(dtype, a->ptr, r, sh, is, ws, 1.0, 0.0, 0, st);
    if (rc) return rc;
    Aptr = ws;
    pa = OperandPlan();
    pa.ok = true; pa.kc = true; pa.ld = P.K > 0 ? P.K : 1; pa.mn = P.M;
  }
  if (P.copy_b) {  // -> [ctr..., free...] contiguous: MN-contiguous, ld = N
    void* wb = (char*)ws + P.ws_a;
    int64_t sh[TNB_MAX_RANK], is[TNB_MAX_RANK];
    int r = 0;
    for (int i = 0; i < nctr; ++i, ++r) { sh[r] = b->shape[P.cb[i]]; is[r] = b->stride[P.cb[i]]; }
    for (int i = 0; i < P.nfb; ++i, ++r) { sh[r] = b->shape[P.fb[i]]; is[r] = b->stride[P.fb[i]]; }
    rc = permute_view(dtype, b->ptr, r, sh, is, wb, 1.0, 0.0, 0, st);
    if (rc) return rc;
    Bptr = wb;
    pb = OperandPlan();
    pb.ok = true; pb.kc = false; pb.ld = P.N > 0 ? P.N : 1; pb.mn = P.N;
  }
  const int opA = pa.kc ? ((cplx && conj_a) ? TNB_OP_J : TNB_OP_N) : ((cplx && conj_a) ? TNB_OP_C : TNB_OP_T);
  const int opB = pb.kc ? ((cplx && conj_b) ? TNB_OP_C : TNB_OP_T) : ((cplx && conj_b) ? TNB_OP_J : TNB_OP_N);
  int64_t batch = 1, sA = 0, sB = 0, sC = 0;
  const int64_t ldc = P.N;
  int64_t Mg = P.M, Ng = P.N;
  if (pa.nbatch > 1) {          // C[(ba, m'), n]
    batch = pa.nbatch; sA = pa.bstride; Mg = pa.mn; sC = pa.mn * P.N;
  } else if (pb.nbatch > 1) {   // C[m, (bb, n')]
    batch = pb.nbatch; sB = pb.bstride; Ng = pb.mn; sC = pb.mn;
  }
  return gemm_ws(dtype, opA, opB, Mg, Ng, P.K, 1.0, 0.0, Aptr, pa.ld, sA, Bptr, pb.ld, sB, 0.0, 0.0, out, ldc, sC, batch,
                 batch == 1 ? sk : nullptr, batch == 1 ? sk_bytes : 0, st);
}
