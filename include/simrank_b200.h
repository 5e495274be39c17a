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
#define SRK_ABI_VERSION 12

int srk_abi_version(void);
const char* srk_last_error(void);
/* Compute capability of the current device as major*10+minor (100 on B200), <0 on error. */
int srk_device_cc(void);

/* Fused-epilogue description shared by the CSR and the tensor-core half-products.
 * With x = the contraction result for output element (r, c):
 *   v = coef * x                      coef carries C (SimRank.py:139) or C1/C2 (:298,:301)
 *   v = v * (1 - 0.5^evidence[r,c])   if evidence != NULL  (SimRank.py:316 with :361/:420/:423)
 *   v = (1-lambda)*v + lambda*prior[r,c]   if prior != NULL (SimRank.py:453,:488,:491)
 *   v = 1 if r + diag_offset == c     np.fill_diagonal(new_S, 1)  (SimRank.py:140,...); diag_offset is the
 *                                     global index of output row 0 (row-sharded CSR calls; the
 *                                     tensor-core calls carry their own diag_offset and ignore this one)
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
  int64_t diag_offset;       /* srk_csr_half_f64 only: see above */
} srk_epilogue;

/* Fixed-point planes (tensor-core path): a non-negative matrix V (R x K) with V[r,k] <= bound(r) is held as
 * NS uint8 planes, plane s at `planes + s*plane_stride`, row-major with leading dimension
 * ldp (multiple of 16), such that  V[r,k] ~= q[r,k] * bound(r) / 256^NS,
 * q = sum_s plane_s[r,k] * 256^(NS-1-s).  The per-row bound is the affine form
 * bound(r) = vec[r]*mul + add (vec == NULL: bound(r) = add), so that the caller can derive it
 * from a constant per-node vector (degree, row sum of G) and per-iteration scalars.          */
typedef struct srk_rowbound {
  const double* vec;
  double mul, add;
} srk_rowbound;

/* ----------------------------------------------------------------------------- CSR path (f64)
 * One half-product  OUT[c, i] = g[i] * sum_{m in N(i)} X[m, c]   (i < M, c < L), i.e.
 * OUT = (G X)^T, written transposed through shared memory so that two calls give
 * (G (G X)^T)^T = G X^T G^T  -- the chain `G.dot(S).dot(G.T)` of SimRank.py:139 (S symmetric).
 * row_begin/row_end restrict i to a row shard (multi-GPU); OUT always has L rows and only its
 * columns [row_begin, row_end) are touched (ldo >= row_end - row_begin: a buffer that holds just
 * those columns is passed as its address minus row_begin elements).
 * final_epi == NULL  -> plain store (first half, T);
 * final_epi != NULL  -> second half with the fused epilogue above (r = c index, c = i).
 * Row-sharded use (one rank owns the output rows [row0, row0 + L) of S_new): X = the column panel
 * T[:, row0 .. row0 + L) as an [K x L] matrix, OUT / s_old / evidence / prior = the rank's row
 * blocks, final_epi->diag_offset = row0.                                                       */
int srk_csr_half_f64(const int64_t* indptr, const int32_t* indices, const double* g,
                     int64_t M, int64_t row_begin, int64_t row_end,
                     const double* X, int64_t ldx, int64_t L,
                     double* OUT, int64_t ldo,
                     const srk_epilogue* final_epi, void* stream);

/* The same half-product with the gather in FIXED POINT (elem = SRK_ELEM_U16), struct-argument form.
 *
 * X is a uint16 matrix with one scale per COLUMN: value(m, c) = X[m, c] * unit(c), unit(c) =
 * in_unit.vec[c] * in_unit.mul + in_unit.add.  The sum over the neighbours m of a graph row is then
 * an exact integer (row degrees < 65536), scaled once per output element -- 4x fewer gathered bytes
 * than float64 with the rounding of the tensor-core path's 2 planes (DESIGN.md "K3").  The unit
 * diagonal of S stays outside the fixed-point matrices (S = I + S_off, as in srk_x2_half):
 *   mode SRK_CSR_FIRST  OUT[c, i] = rint(D[i, c] * unit(c) / (out_bound(i) / qmax)) as uint16, clipped;
 *                       D[i, c] = sum_{m in N(i)} X[m, c].  With X = srk_quantize_rows_u16(S_in) this
 *                       is U = A S_off transposed (`G.dot(S)` of SimRank.py:139 without g), column i
 *                       in units of out_bound(i) / qmax, out_bound(i) >= deg(i) * max(S_off).
 *   mode SRK_CSR_FINAL  x = g[i] * g_col[r] * (D[i, r] * unit(r) + counts[r, i])   (counts = A A^T when
 *                       add_counts, the unit-diagonal term), then the srk_epilogue chain with the
 *                       evidence factor taken from `counts` when use_evidence, stored as float64 at
 *                       OUT[r, i] (r = column of X = output row).
 *   symmetric != 0      (FINAL; square problem, whole row range, no prior, diag_offset 0, OUT row-major
 *                       n x n): only pairs r >= i are computed and every value is stored at (i, r) AND
 *                       (r, i); counts / s_old / evidence are read at (i, r) (they are symmetric).
 * qmax: the sums are exact while deg(i) * qmax < 2^32.  Graphs with a row degree above 65536 (the
 * popular items of a ratings graph) are held with qmax = floor((2^32 - 1) / max degree) < 65535 in
 * every matrix that row gathers from -- the same arithmetic with a coarser step; the units the
 * caller passes (in_unit, and bound / qmax of srk_quantize_rows_u16) follow qmax.
 * elem = SRK_ELEM_F64 is srk_csr_half_f64 (in_unit, out_bound, g_col, counts ignored), plus the
 * symmetric second half.
 * The TMA gather needs 16-byte aligned rows of X (X % 16 == 0, ldx * sizeof(elem) % 16 == 0); other
 * operands are gathered with plain loads (slower, same results).
 *
 * Split neighbour lists (SRK_ELEM_U16; all five fields optional).  A ratings graph has rows with tens of
 * thousands of neighbours (the popular items of BASELINE cfg5) next to rows with a handful: one warp
 * walking such a row is the tail of the whole launch, and the rows of X it gathers from do not stay in L2.
 * The integer sums do not care how a row's list is cut, so the caller may pre-sum pieces of it:
 *   mode SRK_CSR_ACCUM  the "rows" are PIECES of neighbour lists: piece t covers indices[row_lo[t] ..
 *                       row_hi[t]) and its column sums are added (atomically, uint32) to
 *                       accum[accum_slot[t] * ld_accum + c], c < L.  M = number of pieces, the whole
 *                       range [row_begin, row_end) of pieces; indptr, g, OUT are not used.  Pieces
 *                       that are neighbours in t run at the same time: ordering them by the range of X
 *                       rows they gather from keeps that range in L2.
 *   FIRST / FINAL       with row_lo / row_hi given, the list of row r is indices[row_lo[r] .. row_hi[r])
 *                       instead of indptr[r] .. indptr[r + 1]; a row with accum_slot[r] >= 0 starts from
 *                       the sums accum[accum_slot[r] * ld_accum + c] (its pre-summed pieces; the caller
 *                       leaves the rest of its list -- normally nothing -- in row_lo / row_hi).
 *   accum_slot[t] < 0   (ACCUM) piece t is the WHOLE list of its row: its sums are stored (not added) into slot
 *                       -accum_slot[t] - 1, which then needs no zeroing.
 *   mode SRK_CSR_FINISH the second half of SRK_CSR_FINAL alone: every graph row's sums are already in accum
 *                       (slot = row, left there by SRK_CSR_ACCUM over pieces that cover every list); this
 *                       call streams them through the transposed store with the fused epilogue.  Splitting
 *                       the second half this way pays when the rows are short and uneven (a gather CTA of
 *                       16 rows spends a third of its life in the epilogue with its ring idle; the ACCUM
 *                       launch has no epilogue and 8 similar pieces per CTA).  indptr / indices / X are not
 *                       used.
 *                       symmetric != 0: the symmetric second half -- ACCUM was run with symmetric != 0 (a
 *                       piece of row i skips the column panels entirely left of column i) and FINISH
 *                       computes the pairs r >= i from accum[i, r], storing each value at (i, r) and, through a
 *                       shared-memory transposition, at (r, i).
 *   mode SRK_CSR_FINISH_FIRST  the epilogue of SRK_CSR_FIRST alone: OUT[c, i] = rint(accum[i, c] * unit(c) * qmax /
 *                       out_bound(i)) as uint16.
 * accum is zeroed by the caller where pieces are added; ld_accum is a multiple of 512 and >= L rounded up to
 * 512.  Results are bit-identical to the unsplit call: the same integers are added in another order.  */
#define SRK_ELEM_F64 0
#define SRK_ELEM_U16 1
#define SRK_CSR_FIRST 0
#define SRK_CSR_FINAL 1
#define SRK_CSR_ACCUM 2
#define SRK_CSR_FINISH 3
#define SRK_CSR_FINISH_FIRST 4
typedef struct srk_csr_args {
  int elem, mode, symmetric;
  const int64_t* indptr; const int32_t* indices; const double* g;
  int64_t M, row_begin, row_end;
  const void* X; int64_t ldx; int64_t L;
  int64_t K;                                              /* rows of X (= columns of the graph operator); 0 = not given */
  void* OUT; int64_t ldo;
  srk_rowbound in_unit;                                   /* U16: unit of column c of X */
  srk_rowbound out_bound;                                 /* U16 FIRST: bound of output column i */
  double qmax;                                            /* U16: largest fixed-point value, 0 = 65535 (see below) */
  const double* g_col;                                    /* U16 FINAL: row factor of output row r */
  const void* counts; int64_t ld_counts;                  /* uint16 / uint32 A A^T, indexed like OUT */
  int counts_bits, add_counts, use_evidence;
  srk_epilogue epi;                                       /* FINAL */
  const int64_t* row_lo; const int64_t* row_hi;           /* split lists (see above); NULL = indptr */
  uint32_t* accum; int64_t ld_accum;                      /* U16: pre-summed pieces, [slots][ld_accum] */
  const int32_t* accum_slot;                              /* per row (FIRST / FINAL) or per piece (ACCUM) */
} srk_csr_args;
int srk_csr_half(const srk_csr_args* args, void* stream);

/* Source operand of the fixed-point gather: unit[r] = max_k V[r, k] / qmax (element (r, r +
 * zero_diag_offset) excluded and stored as 0; negatives and NaN as 0), XT[k, r] = rint(V[r, k] *
 * (1 / unit[r])) -- the TRANSPOSED matrix [K x ldxt] with one scale per column r.  For the symmetric S of
 * SimRank this is S_off itself with column scales; for a row shard (R local rows) it is the column
 * block the first half gathers from.  Columns R..ldxt-1 are zero-filled.  symmetric != 0: the caller
 * guarantees V == V^T bit for bit (R == K; what the symmetric second half leaves behind), and
 * XT[k, r] is read as V[k, r]: the same result in one streaming pass without the transposition.  */
int srk_quantize_rows_u16(const double* V, int64_t ldv, int64_t R, int64_t K, int64_t zero_diag_offset,
                          uint16_t* XT, int64_t ldxt, double* unit, double qmax, int symmetric, void* stream);

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

/* Edge list -> CSR on the device: what `pivot(index=to, columns=from)` + the row scatter build
 * (SimRank.py:50-52, 199-200), without the dense matrix.  rows / cols are int32 positions of the
 * edge endpoints in the node order the host derived (set order / sorted labels stay host logic),
 * m edges, M rows, K columns.  indptr[M+1] and indices[m] (columns increasing inside each row) are
 * written; *status (device int32) receives bit 0 when a (row, column) pair occurs twice -- the
 * reference's pivot raises "Index contains duplicate entries, cannot reshape" -- and bit 1 when an
 * index is out of range.  workspace: srk_edges_to_csr_workspace(m, M) bytes of device memory.
 * K <= 819200 (a row's columns are sorted through a K-bit bitmap in shared memory).              */
size_t srk_edges_to_csr_workspace(int64_t m, int64_t M);
int srk_edges_to_csr(const int32_t* rows, const int32_t* cols, int64_t m, int64_t M, int64_t K,
                     int64_t* indptr, int32_t* indices, int32_t* status,
                     void* workspace, size_t workspace_bytes, void* stream);

/* ----------------------------------------------------------------------------- tensor-core path */
/* Quantise rows [0,R) of a f64 matrix into planes (round to nearest, clip to 256^NS-1).
 * zero_diag_offset >= 0 forces element (r, r + zero_diag_offset) to 0: the unit diagonal of S
 * is carried separately (S = I + S_off).                                                      */
int srk_slice_rows_f64(const double* V, int64_t ldv, int64_t R, int64_t K,
                       const srk_rowbound* rowbound, int64_t zero_diag_offset, int ns,
                       uint8_t* planes, int64_t ldp, int64_t plane_stride, void* stream);

/* 1 when the tcgen05 path can run on the current device (sm_100), else 0. */
int srk_i8_supported(void);

/* ----------------------------------------------------------------------------- paired-SM tensor-core path
 * The same fixed-point product on a CTA PAIR (tcgen05.mma.cta_group::2, cluster of two SMs):
 *   D[j, r] = sum_k A8[j,k] * V[r,k]      j < M (dense 0/1 matrix), r < R (NS planes), k < K
 * A8 is the M-side operand (256 rows per pair) and the NS planes of RT rows of V are concatenated
 * on the N side of ONE instruction, so the result is row-major in j -- no transposed store between
 * the two half-products -- and the accumulator is double-buffered in TMEM.
 *
 * The unit diagonal of S is carried OUTSIDE the planes (S = I + S_off):
 *   G S G^T = diag(g) (A S_off A^T + A A^T) diag(g),   A A^T = common-in-neighbour counts,
 * a constant uint16 matrix computed once per graph (mode COUNTS).  This keeps every plane bound
 * proportional to max(S_off) instead of 1, which is what lets NS = 2 meet the 1e-6 budget on
 * graphs whose similarities are small (DESIGN.md "precision").
 *
 * The row bounds of U (out_rowbound of MID = in_rowbound of FINAL) are rounded UP to powers of two
 * by the kernel, 2^f >= bound with f >= 8 NS - 40: the caller passes the same srk_rowbound to both.
 *
 * mode SRK_X2_MID   : U[j, r] = D[j,r] * bound_in(r) / 256^NS  re-quantised with 2^f(j) >= out_rowbound(j)
 *                     into out_planes (row j, column r).  With V = planes of S_off this is
 *                     U = A S_off, the first half `G.dot(S)` of SimRank.py:139 without g.
 * mode SRK_X2_FINAL : x = coef * g_a[j] * g_v[r] * (D[j,r] * 2^f(r) / 256^NS + counts), 2^f(r) >= bound_in(r),
 *                     followed by the srk_epilogue chain (the evidence factor is taken from `counts`
 *                     when use_evidence != 0, else from epi.evidence if given), stored as f64.
 *                     layout DIRECT     element (row j, column r) of out_f64/s_old/prior/counts;
 *                     layout SYMMETRIC  M == R, only tiles containing j <= r are computed and every
 *                                       off-diagonal value is written to (j, r) AND (r, j): the second
 *                                       half `.dot(G.T)` costs n^3 instead of 2 n^3 flop and S is
 *                                       bit-exactly symmetric (needs symmetric evidence, no prior);
 *                     layout TRANSPOSED element (row r, column j): row-sharded multi-GPU, where the
 *                                       local rows of S are the N-side operand.
 *                                       With mirror_out != NULL every value is ALSO stored at
 *                                       mirror_out[j * ld_mirror + mirror_col0 + r] -- the mirrored
 *                                       block of the symmetric result, written straight into the
 *                                       memory of the GPU that owns row j (a peer-mapped pointer over
 *                                       NVLink), so no rank computes both (q,p) and (p,q).
 *                     The diagonal is the element with j == r + diag_offset.
 *                     rowmax_hi (optional, caller-zeroed, one uint32 per output row -- row j in the row-major
                     layouts, row r in the transposed one): receives with
 *                     atomicMax a key of the largest off-diagonal value of every row of the result,
 *                     key = high word of the double + 1, so that (double)(key << 32) bounds the row
 *                     within 2^-20: the input of srk_slice_rows_key_f64, which then needs one pass.
 * mode SRK_X2_COUNTS: out_counts[j, r] = min(D[j,r], 65535) as uint16, D[j,r] as uint32 when
 *                     counts_bits == 32 (needed once two rows can share 65535 neighbours), or
 *                     min(D[j,r], 255) as uint8 when counts_bits == 8 (the srk_epilogue.evidence format:
 *                     a count >= 54 already gives exactly 1.0) (ns must be 1, V = a 0/1
 *                     matrix as a single plane): `np.dot((G>0).astype(int), (G>0).T.astype(int))`
 *                     of SimRank.py:315, also the A A^T term above.                              */
#define SRK_X2_MID 0
#define SRK_X2_FINAL 1
#define SRK_X2_COUNTS 2
#define SRK_X2_DIRECT 0
#define SRK_X2_SYMMETRIC 1
#define SRK_X2_TRANSPOSED 2
typedef struct srk_x2_args {
  int mode, ns, layout;
  int64_t M, R, K;
  const uint8_t* A8; int64_t lda;                                    /* [M x K] 0/1 */
  const uint8_t* in_planes; int64_t ld_in; int64_t in_plane_stride;  /* V: NS planes [R x K] */
  /* K-blocked operand (row-sharded multi-GPU exchange buffers): when in_kblock > 0, column k of V is at
   * in_planes + (k / in_kblock) * in_kblock_stride + s * in_plane_stride + r * ld_in + k % in_kblock;
   * in_kblock must be a multiple of 128 and K a multiple of in_kblock.                           */
  int64_t in_kblock; int64_t in_kblock_stride;
  srk_rowbound in_rowbound;                                          /* bound of V row r */
  uint8_t* out_planes; int64_t ld_outp; int64_t out_plane_stride;    /* MID */
  srk_rowbound out_rowbound;                                         /* MID: bound of U row j */
  const double* g_a; const double* g_v;                              /* FINAL: row factors of A8 / V rows */
  const void* counts; int64_t ld_counts;                             /* FINAL (indexed like out_f64) */
  int add_counts, use_evidence;
  int counts_bits;                                                   /* 16 (default when 0) or 32: element type of counts / out_counts; COUNTS also 8 */
  double* out_f64; int64_t ld_out; int64_t diag_offset;              /* FINAL */
  double* mirror_out; int64_t ld_mirror; int64_t mirror_col0;        /* FINAL, TRANSPOSED: see above */
  uint32_t* rowmax_hi;                                               /* FINAL: see above */
  srk_epilogue epi;                                                  /* FINAL */
  void* out_counts; int64_t ld_out_counts;                           /* COUNTS */
  /* Optional scratch (device, 4-byte aligned, contents irrelevant, must not be shared by launches
   * that can overlap): progress counters that keep the CTA pairs of a launch at the same k, so that
   * they find each other's operand panels in L2.  Used when sync_ws_bytes >= 32 * ceil(tiles /
   * CTA pairs) (1 MB covers every supported size); NULL = pairs run free (same results, slower). */
  void* sync_ws; int64_t sync_ws_bytes;
  /* FINAL, TRANSPOSED with mirror_out: like rowmax_hi for the rows of the MIRRORED block, one uint32
   * per A8 row j, updated with atomicMax in the memory of the GPU that owns those rows.        */
  uint32_t* mirror_rowmax_hi;
} srk_x2_args;
int srk_x2_half(const srk_x2_args* args, void* stream);

/* Quantise rows [0,R) of a f64 matrix into NS planes with the EXACT per-row bound: pass 1 takes
 * m_r = max_k V[r,k] over the row (the element (r, r + zero_diag_offset) excluded and stored as
 * 0), pass 2 (served by L2) writes q = rint(V * (256^NS - 1) / m_r).  bound_out[r] receives the
 * bound in the srk_rowbound sense, m_r * 256^NS / (256^NS - 1) (0 for an all-zero row), so the
 * largest element maps to 256^NS - 1 without clipping and the rounding error is at most half a
 * step of m_r / (256^NS - 1).  Negative and NaN entries are stored as 0.                      */
int srk_slice_rows_max_f64(const double* V, int64_t ldv, int64_t R, int64_t K,
                           int64_t zero_diag_offset, int ns,
                           uint8_t* planes, int64_t ldp, int64_t plane_stride,
                           double* bound_out, void* stream);
/* The same planes in ONE pass when the row maxima are already known as keys (rowmax_hi of
 * srk_x2_args): m_r = (double)(key_r << 32) >= max_k V[r,k].  bound_out as above.            */
int srk_slice_rows_key_f64(const double* V, int64_t ldv, int64_t R, int64_t K,
                           int64_t zero_diag_offset, int ns, const uint32_t* rowmax_hi,
                           uint8_t* planes, int64_t ldp, int64_t plane_stride,
                           double* bound_out, void* stream);

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
