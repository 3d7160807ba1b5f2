/*
 * tnb.h -- C ABI of libtnb (tncontract hot path on NVIDIA B200, sm_100a).
 *
 * The reference (andrewdarmawan/tncontract) is pure Python; its "FFI" for the
 * hot path is the NumPy/SciPy call surface listed in SURVEY.md section 8(b).
 * Every entry point below replaces one of those call sites and is what a
 * maintainer would bind (ctypes) in place of the NumPy call -- see
 * INTEGRATION.md for the stubs.  Citations are file:line into
 * /root/reference/tncontract/.
 *
 * Conventions
 *  - all data pointers are DEVICE pointers (cudaMalloc / torch CUDA storage);
 *    the library never frees, retains or reallocates caller buffers;
 *  - `stream` is a cudaStream_t passed as void*; every call is asynchronous
 *    on that stream unless stated otherwise;
 *  - dtype: TNB_F64 (double) or TNB_C128 (interleaved re,im doubles);
 *  - tensors are described by tnb_tensor_t with strides in ELEMENTS;
 *    matrices are row-major with a leading dimension in elements;
 *  - return value: 0 ok, <0 invalid argument (TNB_E_*), >0 a cudaError_t.
 */
#ifndef TNB_H
#define TNB_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TNB_MAX_RANK 24

enum { TNB_F64 = 0, TNB_C128 = 1 };

enum {
  TNB_OK = 0,
  TNB_E_ARG = -1,        /* bad argument (rank, dtype, null pointer, shape mismatch) */
  TNB_E_WORKSPACE = -2,  /* workspace too small */
  TNB_E_UNSUPPORTED = -3,
  TNB_E_NOCONV = -4      /* Jacobi SVD did not converge (maps to numpy.linalg.LinAlgError) */
};

/* operand flags of tnb_gemm */
enum { TNB_OP_N = 0, TNB_OP_T = 1, TNB_OP_C = 2 /* conjugate transpose */, TNB_OP_J = 3 /* conjugate, no transpose */ };

typedef struct {
  void*   ptr;
  int32_t dtype;
  int32_t rank;
  int64_t shape[TNB_MAX_RANK];
  int64_t stride[TNB_MAX_RANK]; /* in elements */
} tnb_tensor_t;

int         tnb_version(void);
const char* tnb_error_string(int code);
/* number of SMs / compute capability of the current device (host query) */
int tnb_device_info(int* sm_count, int* cc_major, int* cc_minor);
/* number of libtnb kernels launched by this process so far (reset != 0 zeroes it) */
long long tnb_launch_count(int reset);

/* Optional device-time profile by kernel class (CUDA events on the launching
 * stream; bench.py uses it for the roofline figures).  enable(1) clears and
 * starts, enable(0) stops.  classes: 0 gemm (work = flops), 1 jacobi rounds
 * (flops executed), 2 qr panel (flops), 3 permute (bytes), 4 mps_mpo_site
 * (bytes), 5 elementwise (bytes).  get() synchronises on the recorded events. */
int tnb_profile_enable(int on);
int tnb_profile_get(int cls, double* ms, double* work, long long* launches, long long* scopes);

/* ---- index permutation: np.rollaxis/np.transpose + np.reshape copy -------
 * replaces the materialised copies behind tensor.py:295,315,354,363,392-394,
 * 482,818,911,1041 and ndarray.copy()/conjugate() (tensor.py:378,485).
 * out (contiguous, row-major, shape[i] = in.shape[perm[i]]) = alpha * op(in),
 * op = conj when `conj` != 0. */
int tnb_permute(const tnb_tensor_t* in, const int32_t* perm, void* out,
                double alpha_re, double alpha_im, int conj, void* stream);

/* ---- in-place / elementwise helpers ------------------ */
/* x *= alpha on an arbitrary strided view: onedim_core.py:280,311,313,358,479,482 */
int tnb_scale_inplace(const tnb_tensor_t* x, double alpha_re, double alpha_im, void* stream);
/* out(contiguous) = a*x + b*y, same logical shape: tensor.py:137, :800 */
int tnb_axpby(const tnb_tensor_t* x, const tnb_tensor_t* y, void* out,
              double a_re, double a_im, double b_re, double b_im, void* stream);
/* Frobenius norm -> *out_device (double): np.linalg.norm, tensor.py:590,
 * onedim_core.py:272,305,311,656,658.  ws: >= tnb_norm2_workspace() bytes. */
size_t tnb_norm2_workspace(void);
int tnb_norm2(const tnb_tensor_t* x, double* out_device, void* ws, size_t ws_bytes, void* stream);
/* np.diag both ways: tensor.py:940,1133,1158; onedim_core.py:321,340.
 * embed: out is an n x n contiguous matrix of `dtype`, s is a real vector;
 * mode 0: s, 1: sqrt(s) (tensor.py:954,1178), 2: 1/s (inv of a diagonal, tensor.py:488). */
int tnb_diag_embed(int dtype, const double* s, int64_t n, void* out, int mode, void* stream);
/* extract the diagonal of a 2-D view into a contiguous vector of the same dtype */
int tnb_diag_extract(const tnb_tensor_t* x, void* out, void* stream);
/* rows (axis 0) of a contiguous [n, cols] matrix scaled by f(s[i]), mode as
 * above: the dense-diagonal contractions onedim_core.py:349, tensor.py:945-956,
 * 1171-1180 without the k x k GEMM.  axis = 0 scales rows, 1 scales columns. */
int tnb_diag_scale(int dtype, void* x, int64_t rows, int64_t cols, int64_t ld, const double* s,
                   int axis, int mode, void* stream);
/* np.trace over two axes of a view (tensor.py:329): out contiguous, remaining axes in order */
int tnb_trace(const tnb_tensor_t* x, int axis1, int axis2, void* out, void* stream);
/* widen a real vector/tensor to complex128 (dtype promotion inside np.tensordot) */
int tnb_real_to_complex(const double* x, int64_t n, void* out, void* stream);

/* out[i] = U[0,1) from a counter-based generator (splitmix64 of key + (offset+i)*golden):
 * stands in for np.random.rand (onedim_utils.py:47) when the batched path creates its
 * synthetic networks on the owning GPU; reproducible element-wise on the host. */
int tnb_fill_uniform(double* out, int64_t n, unsigned long long key, unsigned long long offset, void* stream);

/* ---- GEMM: the BLAS call under np.tensordot (tensor.py:735) ---------------
 * C[b] = alpha * op(A[b]) * op(B[b]) + beta * C[b], row-major, b < batch,
 * X[b] = X + b*strideX (elements).  FP64 / complex128 on DMMA tensor cores. */
int tnb_gemm(int dtype, int opA, int opB, int64_t M, int64_t N, int64_t K,
             const double* alpha /*[2]*/, const void* A, int64_t lda, int64_t strideA,
             const void* B, int64_t ldb, int64_t strideB,
             const double* beta /*[2]*/, void* C, int64_t ldc, int64_t strideC,
             int64_t batch, void* stream);
/* The same product with caller-provided scratch for split-K: a product with few C tiles and a long K
 * (Gram-like shapes: the Q^H W products under np.linalg.qr, tensor.py:1044) is cut along K over
 * grid.z and reduced in a second, deterministic kernel.  Any ws_bytes >= 0 is valid (more scratch
 * allows more splits; M*N*sizeof(elem) per split); batch must be 1 for the split to apply. */
int tnb_gemm_ws(int dtype, int opA, int opB, int64_t M, int64_t N, int64_t K,
                const double* alpha /*[2]*/, const void* A, int64_t lda, const void* B, int64_t ldb,
                const double* beta /*[2]*/, void* C, int64_t ldc, void* ws, size_t ws_bytes, void* stream);

/* ---- tensordot: np.tensordot(a, b, (axes_a, axes_b)) (tensor.py:735) -------
 * out is contiguous with shape free(a) ++ free(b).  conj flags fold
 * Tensor.conjugate() (tensor.py:484) into the operand load.  Operand
 * permutations are fused into GEMM operand staging whenever the strides
 * allow (N / T / strided-batched forms); otherwise the operand is permuted
 * into `ws` first.  ws must hold tnb_tensordot_workspace() bytes. */
size_t tnb_tensordot_workspace(const tnb_tensor_t* a, const tnb_tensor_t* b, int nctr,
                               const int32_t* axes_a, const int32_t* axes_b);
int tnb_tensordot(const tnb_tensor_t* a, const tnb_tensor_t* b, int nctr,
                  const int32_t* axes_a, const int32_t* axes_b, int conj_a, int conj_b,
                  void* out, void* ws, size_t ws_bytes, void* stream);

/* ---- MPO x MPS site apply: onedim_core.py:1702-1704 ------------------------
 * fuses tensordot(A[phys,left,right], W[left,right,physout,physin]) over phys
 * with consolidate_indices(): out[(l,wl), physout, (r,wr)] contiguous.
 * A: (d, Dl, Dr) view, W: (wl, wr, dout, d) view (any strides). */
int tnb_mps_mpo_site(const tnb_tensor_t* A, const tnb_tensor_t* W, void* out, void* stream);

/* ---- QR: np.linalg.qr(mode="reduced") (tensor.py:1044) ---------------------
 * A: m x n row-major (lda); Q: m x k, R: k x n contiguous, k = min(m,n).
 * k >= 128: block Gram-Schmidt with reorthogonalisation (BCGS-PIP+: every O(m n^2) step a DMMA
 * GEMM, 64 x 64 Cholesky factors, device-side breakdown checks; the second pass runs per group of 256
 * columns and is skipped on the device for a group its first pass left orthonormal to 1e-14 entrywise); k < 128 and as the fallback of
 * the former: blocked Householder (compact WY, cluster panels).  R has a real diagonal -- positive
 * on the Gram-Schmidt path, LAPACK geqrf's sign convention on the Householder path; Q R = A and
 * Q^H Q = I either way (callers on the path only use the product and the isometry).
 * ws: tnb_qr_workspace() bytes.  Synchronises `stream` once for k >= 128 (fallback flag). */
size_t tnb_qr_workspace(int dtype, int64_t m, int64_t n);
int tnb_qr(int dtype, int64_t m, int64_t n, const void* A, int64_t lda,
           void* Q, void* R, void* ws, size_t ws_bytes, void* stream);

/* ---- SVD: np.linalg.svd(full_matrices=False) (tensor.py:915) ---------------
 * A: m x n row-major (lda); U: m x k, S: k doubles (descending), Vh: k x n,
 * k = min(m,n).  QR/LQ pre-reduction + block one-sided Jacobi.  Synchronises
 * the stream once per batch of queued Jacobi sweeps (convergence flag; the first batch is as long as the
 * previous factorisation needed, so normally once per call).  Returns TNB_E_NOCONV
 * if max_sweeps is exhausted or the input holds NaN / Inf. */
size_t tnb_svd_workspace(int dtype, int64_t m, int64_t n);
int tnb_svd(int dtype, int64_t m, int64_t n, const void* A, int64_t lda,
            void* U, double* S, void* Vh, void* ws, size_t ws_bytes, int* sweeps_out, void* stream);

/* Projection form used by the MPS sweeps (onedim_core.py:317-351): U and S as above, and
 * P = U^H A = diag(S) Vh (k x n) -- the product the sweep absorbs into the next site.  V is never
 * accumulated (one third fewer Jacobi flops); the rows of P carry an absolute error of eps |A|.
 * A wide matrix (m < n) takes U from Jacobi on R^H of A^H = Q R; a tall or square one factors R^H = Q2 R2 once
 * more and runs on R2^H, which has the left vectors of R (Jacobi on the columns of R itself needs 2-4 x the sweeps
 * on rank-deficient input).  U is an isometry over all k columns, noise directions included.
 * Same workspace as tnb_svd. */
int tnb_svd_project(int dtype, int64_t m, int64_t n, const void* A, int64_t lda,
                    void* U, double* S, void* P, void* ws, size_t ws_bytes, int* sweeps_out, void* stream);

/* ---- truncation rule on device ---------------------------------------------
 * kept = #{ i < (chi>0 ? chi : n) : s[i] > bar }, bar = threshold (relative
 * == 0, absolute) or threshold*s[0] (relative == 1, tensor.py:1150); relative
 * == 2 tests s[i]/s[0] > threshold instead (the form of onedim_core.py:333-336);
 * tensor.py:1137-1152 (chi first, then
 * threshold) and onedim_core.py:333-339 (relative to s[0], then chi) give the
 * same count because s is sorted.  info_device[0] = kept (as a double),
 * info_device[1] = s[0], so the host learns both with one 16-byte read;
 * if s_scaled != NULL it receives s / s[0] (onedim_core.py:333). */
int tnb_truncation_count(const double* s, int64_t n, int64_t chi, double threshold, int relative,
                         double* info_device, double* s_scaled, void* stream);

/* ==== batched path (BASELINE.json config 4: many independent networks of one shape) ==================
 * The reference runs one network per Python loop iteration (onedim_core.py:463-484 svd_compress,
 * :1666-1683 inner_product_mps); here ONE launch serves the same site of every network of a shard.
 * Matrix b of a batch lives at base + b * stride (elements); all matrices of a batch share one shape.
 * tnb_gemm above is already strided-batched and serves the absorbs and the ladder contractions. */

/* batch x tnb_svd_project (np.linalg.svd, tensor.py:915, as used at onedim_core.py:317-351):
 * A[b]: m x n (lda) -> U[b]: m x k, S[b]: k doubles (descending), P[b] = U^H A = diag(S) Vh: k x n (may be
 * NULL), k = min(m, n).  One-sided Jacobi directly on the columns of A (tall) or A^H (wide), no QR
 * pre-reduction; grid row = matrix, every matrix has its own convergence flag; the stream is synchronised
 * once per batch of queued sweeps.  sweeps_out: the largest sweep count of the batch. */
size_t tnb_svd_project_batched_workspace(int dtype, int64_t m, int64_t n, int64_t batch);
int tnb_svd_project_batched(int dtype, int64_t m, int64_t n, int64_t batch, const void* A, int64_t lda, int64_t strideA,
                            void* U, int64_t strideU, double* S, int64_t strideS, void* P, int64_t strideP,
                            void* ws, size_t ws_bytes, int* sweeps_out, void* stream);
/* batch x tnb_truncation_count: info_device[2b] = kept, [2b+1] = s[b][0]; s_scaled (stride as s) optional */
int tnb_truncation_count_batched(const double* s, int64_t n, int64_t stride, int64_t batch, int64_t chi,
                                 double threshold, int relative, double* info_device, double* s_scaled, void* stream);
/* out_device[b] = Frobenius norm of the `per` contiguous elements at x + b * stride (np.linalg.norm,
 * onedim_core.py:656-658 norm(canonical_form=...)) */
int tnb_norm2_batched(int dtype, const void* x, int64_t per, int64_t stride, int64_t batch, double* out_device, void* stream);
/* batch x tnb_fill_uniform: matrix b (per contiguous doubles) is filled from stream keys_device[b] */
int tnb_fill_uniform_batched(double* out, int64_t per, int64_t batch, const unsigned long long* keys_device, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* TNB_H */
