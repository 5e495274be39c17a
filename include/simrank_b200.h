/* simrank_b200.h -- C ABI of the B200-native SimRank iteration engine (libsimrank_b200.so).
 *
 * The reference (ysong1231/SimRank) is pure Python with no FFI of its own; the drop-in
 * boundary is its class API (SimRank/SimRank.py).  This header is the native boundary a
 * maintainer binds from that Python code (ctypes stub in INTEGRATION.md).  Each entry point
 * names the reference expression it replaces (file:line, relative to the reference repo).
 *
 * Conventions
 *   - every function returns 0 on success, <0 on error; srk_last_error() gives the text;
 *   - the CALLER owns every buffer (device pointers unless the name says host); the library
 *     allocates nothing that outlives a call;
 *   - every call is asynchronous on the cudaStream_t passed as `void* stream`;
 *   - matrices are row-major with an explicit leading dimension in ELEMENTS;
 *   - a graph operator is a row-scaled 0/1 matrix G = diag(g) * A.  The reference's
 *     _create_graph can only produce such matrices (SimRank.py:49,197-198: the value of every
 *     in-edge of a node is 1/inNeighbors(node)); SimRank++ weights W = diag(spread)*G
 *     (SimRank.py:333) keep that form with g' = spread*g.
 *   - A is given either as CSR (int64 indptr[M+1], int32 indices[nnz], column indices sorted
 *     within a row) or as a dense uint8 0/1 matrix for the tensor-core path.
 */
#ifndef SIMRANK_B200_H_
#define SIMRANK_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SRK_OK 0
#define SRK_ERR_INVALID (-1)
#define SRK_ERR_CUDA (-2)
#define SRK_ERR_UNSUPPORTED (-3)

/* ABI version of this header; srk_abi_version() must return the same value. */
#define SRK_ABI_VERSION 1

int srk_abi_version(void);
const char* srk_last_error(void);
/* Compute capability of the current device as major*10+minor (100 on B200), <0 on error. */
int srk_device_cc(void);

/* Fused-epilogue description shared by the CSR and the tensor-core half-products.
 * With x = the contraction result for output element (r, c):
 *   v = coef * x                      coef carries C (SimRank.py:139) or C1/C2 (:298,:301)
 *   v = v * (1 - 0.5^evidence[r,c])   if evidence != NULL  (SimRank.py:316 with :361/:420/:423)
 *   v = (1-lambda)*v + lambda*prior[r,c]   if prior != NULL (SimRank.py:453,:488,:491)
 *   v = 1 if r == c                   np.fill_diagonal(new_S, 1)  (SimRank.py:140,...)
 *   *maxdiff = max(*maxdiff, |v - s_old[r,c]|)   if s_old != NULL: the reduction behind
 *                                     _converged (SimRank.py:74): converged <=> maxdiff <= eps
 *   *maxoff  = max(*maxoff, v) over r != c        if maxoff != NULL (range tracking for the
 *                                     fixed-point slices of the tensor-core path)
 * maxdiff / maxoff are device doubles the caller zeroes before the call; NaN differences are
 * ignored, as in the reference where (abs(a-b) > eps) is False for NaN.                    */
typedef struct srk_epilogue {
  double coef;
  const uint8_t* evidence;   /* common-neighbour counts clipped to 255 (count>=54 == 1.0), or NULL */
  int64_t ld_evidence;
  const double* prior;       /* or NULL */
  int64_t ld_prior;
  double lambda;
  const double* s_old;       /* or NULL */
  int64_t ld_s_old;
  double* maxdiff;           /* device scalar or NULL */
  double* maxoff;            /* device scalar or NULL */
} srk_epilogue;

/* ----------------------------------------------------------------------------- CSR path (f64)
 * One half-product  OUT[c, i] = g[i] * sum_{m in N(i)} X[m, c]   (i < M, c < L), i.e.
 * OUT = (G X)^T, written transposed through shared memory so that two calls give
 * (G (G X)^T)^T = G X^T G^T  -- the chain `G.dot(S).dot(G.T)` of SimRank.py:139 (S symmetric).
 * row_begin/row_end restrict i to a row shard (multi-GPU); OUT always has L rows.
 * final_epi == NULL  -> plain store (first half, T);
 * final_epi != NULL  -> second half with the fused epilogue above (r = c index, c = i).     */
int srk_csr_half_f64(const int64_t* indptr, const int32_t* indices, const double* g,
                     int64_t M, int64_t row_begin, int64_t row_end,
                     const double* X, int64_t ldx, int64_t L,
                     double* OUT, int64_t ldo,
                     const srk_epilogue* final_epi, void* stream);

/* Common-in-neighbour counts cnt[i,j] = |N(i) & N(j)| clipped to 255, for rows
 * [row_begin,row_end) x all j < M: `np.dot((G>0).astype(int), (G>0).T.astype(int))`
 * (SimRank.py:315).  Rows flagged in `dead` (g<=0: G>0 is False there) count as empty.      */
int srk_csr_evidence_counts(const int64_t* indptr, const int32_t* indices, const uint8_t* dead,
                            int64_t M, int64_t row_begin, int64_t row_end,
                            uint8_t* counts, int64_t ldc, void* stream);

/* spread[i] = exp(-var_i), var_i = sample variance (ddof=1) of the nonzero values of row i,
 * NaN (fewer than 2 nonzeros) -> 0:  `G.replace(0,nan).var(axis=1).fillna(0).apply(exp(-x))`
 * (SimRank.py:326-332).  vals == NULL means every stored entry of row i equals g[i].         */
int srk_csr_row_spread(const int64_t* indptr, const double* vals, const double* g, int64_t M,
                       double* spread, void* stream);

/* Scatter a CSR 0/1 pattern into a dense uint8 matrix (rows [row_begin,row_end) -> A8 rows
 * 0..), zero-filling the rest; replaces the pivot + row scatter of SimRank.py:50-52.          */
int srk_csr_to_dense_u8(const int64_t* indptr, const int32_t* indices, int64_t row_begin,
                        int64_t row_end, int64_t K, uint8_t* A8, int64_t lda, void* stream);

/* ----------------------------------------------------------------------------- tensor-core path
 * Fixed-point planes: a non-negative matrix V (R x K) with V[r,k] <= bound(r) is held as
 * NS uint8 planes, plane s at `planes + s*plane_stride`, row-major with leading dimension
 * ldp (multiple of 16), such that  V[r,k] ~= q[r,k] * bound(r) / 256^NS,
 * q = sum_s plane_s[r,k] * 256^(NS-1-s).  The per-row bound is the affine form
 * bound(r) = vec[r]*mul + add (vec == NULL: bound(r) = add), so that the caller can derive it
 * from a constant per-node vector (degree, row sum of G) and per-iteration scalars.          */
typedef struct srk_rowbound {
  const double* vec;
  double mul, add;
} srk_rowbound;

/* Quantise rows [0,R) of a f64 matrix into planes (round to nearest, clip to 256^NS-1).
 * zero_diag_offset >= 0 forces element (r, r + zero_diag_offset) to 0: the unit diagonal of S
 * is carried separately (S = I + S_off).                                                      */
int srk_slice_rows_f64(const double* V, int64_t ldv, int64_t R, int64_t K,
                       const srk_rowbound* rowbound, int64_t zero_diag_offset, int ns,
                       uint8_t* planes, int64_t ldp, int64_t plane_stride, void* stream);

/* Workspace-free tensor-core half-product (tcgen05.mma kind::i8, TMEM accumulators, TMA feed):
 *   D[r, j] = sum_k V[r,k] * A8[j,k]      r < R (planes), j < N (dense 0/1 matrix, K columns)
 * mode SRK_I8_MID   : U[j, r] = (D[r,j] + unit_diag * A8[j, r + diag_offset]) is re-quantised
 *                     into `out_planes` (row j, column r: TRANSPOSED store) with bound
 *                     out_rowbound[j]; this is the first half `G.dot(S)` of SimRank.py:139.
 * mode SRK_I8_FINAL : S_new[r, j] = epilogue(g_row[r] * g_col[j] * D[r,j]) stored as f64 into
 *                     `out_f64` (+ optionally re-quantised into out_planes with out_rowbound[r],
 *                     diagonal forced to 0): the second half `.dot(G.T)` with everything of
 *                     SimRank.py:138-140 and :74 fused.
 * mode SRK_I8_COUNTS: counts[r, j] = min(D[r,j], 255) as uint8 into out_planes (ns must be 1,
 *                     in_rowbound ignored): the evidence product of SimRank.py:315.
 * diag_offset is the global row index of local row 0 (row-sharded operands).                */
#define SRK_I8_MID 0
#define SRK_I8_FINAL 1
#define SRK_I8_COUNTS 2
typedef struct srk_i8_args {
  int mode, ns;
  int64_t R, N, K;
  const uint8_t* in_planes; int64_t ld_in; int64_t in_plane_stride;
  /* K-blocked operand (row-sharded multi-GPU exchange buffers): when in_kblock > 0, column k of
   * V is at in_planes + (k / in_kblock) * in_kblock_stride + s * in_plane_stride + r * ld_in +
   * k % in_kblock; in_kblock must be a multiple of 128 and K a multiple of in_kblock.          */
  int64_t in_kblock; int64_t in_kblock_stride;
  srk_rowbound in_rowbound;             /* bound of V row r */
  const uint8_t* A8; int64_t lda;       /* [N x K] 0/1 */
  int64_t diag_offset; int unit_diag;
  const double* g_row; const double* g_col;          /* FINAL */
  double* out_f64; int64_t ld_out;                   /* FINAL */
  uint8_t* out_planes; int64_t ld_outp; int64_t out_plane_stride;
  srk_rowbound out_rowbound;            /* MID: bound of U row j; FINAL: bound of S_new row r */
  srk_epilogue epi;                                  /* FINAL */
} srk_i8_args;
int srk_i8_half(const srk_i8_args* args, void* stream);
/* 1 when the tcgen05 path can run on the current device (sm_100), else 0. */
int srk_i8_supported(void);

/* ----------------------------------------------------------------------------- retrieval
 * Row-wise top-k of S (rows [0,R) x n columns): idx[r,:] = argsort(-S[r], kind='stable')[:k]
 * (ties -> lower column first; NaN last), vals the matching values.  No reference call site:
 * the reference returns whole DataFrames (SimRank.py:141); oracle = oracle.topk.              */
int srk_topk_rows(const double* S, int64_t lds, int64_t R, int64_t n, int k,
                  int32_t* idx, double* vals, void* stream);

/* S <- I (R x n block whose global first row is diag_offset).  SimRank.py:124-126.          */
int srk_set_identity_f64(double* S, int64_t lds, int64_t R, int64_t n, int64_t diag_offset,
                         void* stream);

#ifdef __cplusplus
}
#endif
#endif /* SIMRANK_B200_H_ */
