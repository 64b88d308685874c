/*
 * sgp_b200 — C ABI of the B200 (sm_100a) kernels behind the SGP training-free encoder.
 *
 * The reference (Graph-Machine-Learning-Group/sgp) is pure Python and has no FFI of its own:
 * its "plugin" surface is the Python classes/functions of lib/nn/encoders/*, lib/sgp_preprocessing.py
 * and lib/nn/reservoir/reservoir.py.  Every entry point below replaces the arithmetic of one of
 * those call sites (cited per function, paths relative to the reference root) and is what a
 * maintainer would bind with ctypes from those files (INTEGRATION.md shows the stubs).
 *
 * Conventions
 *   - plain C: raw device pointers, sizes, strides (in ELEMENTS), a cudaStream_t passed as void*;
 *     no torch types.
 *   - every function returns 0 on success, a negative SGP_E* code otherwise; it never throws,
 *     never allocates (the caller owns every buffer, workspaces are sized by *_workspace_bytes),
 *     is stream-ordered on `stream`, and is re-entrant across streams.
 *   - sgp_last_error() returns a thread-local human-readable message for the last failure.
 *   - all floating point is IEEE fp32 (the reference's precision=32), indices are int32 in the
 *     CSR (the reference's int64 edge_index is narrowed while building it; N < 2^31, nnz < 2^31).
 *   - there is no CPU fallback anywhere behind this ABI.
 */
#ifndef SGP_B200_H
#define SGP_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* bumped with every change of a signature or of a kernel's contract; sgp_version() returns it and
 * the Python binding refuses a library built from another header */
#define SGP_B200_ABI_VERSION 200

#define SGP_OK 0
#define SGP_EINVAL (-1)    /* bad shape / flag / null pointer */
#define SGP_EALIGN (-2)    /* pointer or stride not aligned as the kernel requires */
#define SGP_ECUDA (-3)     /* CUDA runtime error (message in sgp_last_error) */
#define SGP_ECAPACITY (-4) /* caller-provided buffer too small */
#define SGP_EUNSUPPORTED (-5)

/* activation codes (reference: lib/nn/reservoir/reservoir.py:37-41) */
#define SGP_ACT_TANH 0
#define SGP_ACT_RELU 1
#define SGP_ACT_SELF_NORM 2
#define SGP_ACT_IDENTITY 3

/* sgp_csr_build flags (reference: lib/sgp_preprocessing.py:67-105, 182-185) */
#define SGP_CSR_SET_DIAG 1     /* set_diag(): drop stored diagonal, insert (i,i)=1 for all i */
#define SGP_CSR_REMOVE_DIAG 2  /* remove_diag(): drop stored diagonal (ignored when SET_DIAG) */
#define SGP_CSR_GCN_NORM 4     /* D^-1/2 S D^-1/2 instead of D^-1 S */
#define SGP_CSR_SYMMETRIZE 8   /* to_undirected(): add reversed edges, coalesce duplicates by add */
#define SGP_CSR_TRANSPOSE 16   /* swap the two rows of edge_index first (the bidirectional pass) */
#define SGP_CSR_NO_NORM 32     /* keep the weights as given (an already normalised adjacency: GESNLayer.forward) */

int sgp_version(void);
const char* sgp_last_error(void);
/* Number of kernels this library has launched in the calling process (for bench accounting). */
int64_t sgp_launch_count(void);

/* ---------------------------------------------------------------------------------------------
 * K3  adjacency -> normalised CSR.
 * Replaces preprocess_adj (lib/sgp_preprocessing.py:67-105) and the to_undirected /
 * edge_index[[1,0]] handling of sgp_spatial_embedding (:182-185, :205-207).
 *   edge_src = edge_index[0] (the COLUMN / source j), edge_dst = edge_index[1] (the ROW / target i)
 *   (":80  col, row = edge_index").  Entries are ordered by (row, col), duplicates are kept
 *   (and therefore summed by the SpMM) unless SYMMETRIZE coalesces them.
 *   weight may be NULL (unit weights).  cap >= 2*E + N is always enough.
 *   *nnz_out (HOST pointer) receives the number of stored entries; the call synchronises `stream`.
 * ------------------------------------------------------------------------------------------- */
size_t sgp_csr_build_workspace_bytes(int64_t E, int32_t N, int flags);
int sgp_csr_build(const int64_t* edge_src, const int64_t* edge_dst, const float* weight,
                  int64_t E, int32_t N, int flags,
                  int32_t* rowptr /*[N+1]*/, int32_t* col /*[cap]*/, float* val /*[cap]*/,
                  int64_t cap, int64_t* nnz_out,
                  void* workspace, size_t workspace_bytes, void* stream);

/* ---------------------------------------------------------------------------------------------
 * K1  fused leaky-ESN scan of one reservoir layer over a chunk of Tc time steps.
 * Replaces the Python time loop Reservoir.forward (lib/nn/reservoir/reservoir.py:158-186) around
 * ReservoirLayer.forward (:77-81):
 *     h' = (1-alpha) h + alpha * act( x_t W_ih^T + b + h W_hh^T )
 * for every node n and step t of the chunk; h' is written to out[t, n, 0:H] and carried.
 *   wpack   [sgp_reservoir_pack_rows(Fin, H), H] = [(FinP + H), H], FinP = Fin zero-padded to the
 *           kernel's k-tile: rows 0..Fin-1 = W_ih^T, rows Fin..FinP-1 = 0, rows FinP.. = W_hh^T
 *           (made by sgp_reservoir_pack)
 *   h_state [N, H] in/out: state before the first / after the last step of the chunk
 *   x       element (t, n, f) at x[t*x_t_stride + n*x_n_stride + f]
 *   out     element (t, n, j) at out[t*out_t_stride + n*out_n_stride + j]  (a feature block of
 *           the concatenated [Tc, N, D] encoder output; deeper layers read the previous layer's
 *           block as their x with Fin = H)
 * H in {128, 256} with 16-byte aligned out / h_state / wpack and strides % 4 == 0 runs the tiled
 * FFMA2 kernel; every other shape runs the generic (still CUDA) kernel.
 * ------------------------------------------------------------------------------------------- */
int sgp_reservoir_pack_rows(int Fin, int H);
int sgp_reservoir_pack(const float* w_ih /*[H,Fin]*/, const float* w_hh /*[H,H]*/, int Fin, int H,
                       float* wpack /*[(FinP+H),H]*/, void* stream);
int sgp_reservoir_scan(const float* x, int64_t x_t_stride, int64_t x_n_stride, int Fin,
                       const float* wpack, const float* bias /*[H]*/,
                       float alpha, float one_minus_alpha, int act,
                       float* h_state,
                       float* out, int64_t out_t_stride, int64_t out_n_stride,
                       int Tc, int N, int H, void* stream);

/* All layers of a SMALL reservoir in one launch (H in {16, 32, 64}, L <= 8, Fin <= 64): the shipped
 * configurations of the reference (config/traffic/sgp_la.yaml H = 64 x L = 2,
 * config/largescale_100nn/sgp_pv.yaml H = 16 x L = 8).  Layer l reads layer l-1's new state of the
 * same step from shared memory (reservoir.py:170-176) and all weights stay resident there.
 * w_ih / w_hh / bias: HOST arrays of L device pointers ([H, Fin_l] with Fin_0 = Fin, Fin_l = H;
 * [H, H]; [H], exactly the reference's parameters — no packing); alpha: HOST array [L];
 * h_state [L, N, H] in/out; out element (t, n, l*H + j).  SGP_EUNSUPPORTED when the shape does not
 * fit (the caller then runs sgp_reservoir_scan layer by layer). */
int sgp_reservoir_scan_multi(const float* x, int64_t x_t_stride, int64_t x_n_stride, int Fin,
                             const float* const* w_ih, const float* const* w_hh, const float* const* bias,
                             const float* alpha, int act, float* h_state,
                             float* out, int64_t out_t_stride, int64_t out_n_stride,
                             int Tc, int N, int H, int L, void* stream);

/* Tensor-core scan (tcgen05, 3xTF32, fp32-accurate): same contract as sgp_reservoir_scan for
 * H in {128, 256}, Fin <= 8, activation tanh / relu / identity.  wimg [H*H*2] = W_hh split into tf32
 * hi / lo images in the kernel's shared-memory layout (sgp_reservoir_tc_pack); w_ih [H, Fin] and bias
 * [H] as in the reference.  *err_flag (device int) is set to 1 if an internal barrier times out.
 * checksum (device fp64 scalar, may be NULL): the kernel adds the sum of every state value it
 * writes to out — the streamed benchmark's sink, accumulated in the epilogue registers instead of
 * re-reading the output (sgp_checksum). */
int sgp_reservoir_tc_pack(const float* w_hh /*[H,H]*/, int H, float* wimg /*[2*H*H]*/, void* stream);
int sgp_reservoir_scan_tc(const float* x, int64_t x_t_stride, int64_t x_n_stride, int Fin,
                          const float* wimg, const float* w_ih, const float* bias,
                          float alpha, float one_minus_alpha, int act,
                          float* h_state, float* out, int64_t out_t_stride, int64_t out_n_stride,
                          int Tc, int N, int H, int* err_flag, double* checksum /*nullable*/, void* stream);

/* Tensor-core scan with fp16x3 operands (tcgen05 kind::f16: twice the tf32 MMA rate, half the W
 * stream; same 22-bit split accuracy): same contract as sgp_reservoir_scan_tc for TANH reservoirs
 * whose states stay within [-1, 1] (zero or bounded initial state: tanh + leaky blend keep them
 * there).  w_scale = the power of two that brings max|W_hh| into [2^13, 2^14) (fp16 normal range;
 * chosen by the caller, passed to both calls); wimg [2*H*H] fp16 (sgp_reservoir_tc16_pack). */
int sgp_reservoir_tc16_pack(const float* w_hh /*[H,H]*/, int H, float w_scale, void* wimg /*fp16 [2*H*H]*/,
                            void* stream);
int sgp_reservoir_scan_tc16(const float* x, int64_t x_t_stride, int64_t x_n_stride, int Fin,
                            const void* wimg, float w_scale, const float* w_ih, const float* bias,
                            float alpha, float one_minus_alpha,
                            float* h_state, float* out, int64_t out_t_stride, int64_t out_n_stride,
                            int Tc, int N, int H, int* err_flag, double* checksum /*nullable*/, void* stream);

/* ---------------------------------------------------------------------------------------------
 * K2  CSR x dense propagation, batched over the leading (time / batch) axis.
 * Replaces `x = adj @ x` (lib/sgp_preprocessing.py:200-203; torch_sparse spmm_sum):
 *     dst[t, i, :] = sum_{e in row i} val[e] * src[t, col[e], :]      i in [0, n_rows)
 *   row_order  optional [n_rows] permutation: the order in which rows are SCHEDULED (locality),
 *              results are still written at row i.  NULL = natural order.
 *   src/dst    element (t, n, f) at p[t*t_stride + n*n_stride + f]; src and dst must not overlap.
 * sgp_khop_spmm runs `hops` successive hops inside one [Tc, N, D] buffer: hop h reads feature
 * block (h == 0 ? block_in : block_out0 + h - 1) and writes block block_out0 + h, i.e. it fills
 * the reference's `res` list (:200-203) directly in its concatenated layout
 * (lib/nn/encoders/sgp_spatial_encoder.py:35), without the torch.cat copy.
 * ------------------------------------------------------------------------------------------- */
int sgp_spmm(const int32_t* rowptr, const int32_t* col, const float* val, const int32_t* row_order,
             const float* src, int64_t src_t_stride, int64_t src_n_stride,
             float* dst, int64_t dst_t_stride, int64_t dst_n_stride,
             int n_rows, int F, int Tc, void* stream);
/* Row-sharded variant: column ids < n_split address `src` (the rank's own rows), ids >= n_split
 * address row (id - n_split) of `src2` (halo rows received from the other ranks). */
int sgp_spmm_halo(const int32_t* rowptr, const int32_t* col, const float* val, const int32_t* row_order,
                  const float* src, int64_t src_t_stride, int64_t src_n_stride,
                  const float* src2, int64_t src2_t_stride, int64_t src2_n_stride, int n_split,
                  float* dst, int64_t dst_t_stride, int64_t dst_n_stride,
                  int n_rows, int F, int Tc, void* stream);
int sgp_khop_spmm(const int32_t* rowptr, const int32_t* col, const float* val,
                  const int32_t* row_order,
                  float* buf, int64_t t_stride, int64_t n_stride,
                  int block_in, int block_out0, int hops, int N, int F, int Tc, void* stream);

/* Row-block-union (RBU) operator: rows are grouped R at a time (R in {4,8,16}); each group stores
 * the sorted union of its column indices once plus a dense [U, R] value slab, so that every
 * gathered source row is loaded once per group and reused from registers for R output rows.
 *   grp_ptr  [n_groups+1] offsets into ucol / uval
 *   grp_rows [n_groups, R] destination row of each slot (-1 = padding)
 *   ucol     [total_U]     source row ids
 *   uval     [total_U, R]  operator values (0 where the row does not have that column)
 * Requires F % 128 == 0.  Same result as sgp_spmm on the CSR the groups were built from. */
int sgp_spmm_rbu(const int32_t* grp_ptr, const int32_t* grp_rows, const int32_t* ucol,
                 const float* uval, int R, int n_groups,
                 const float* src, int64_t src_t_stride, int64_t src_n_stride,
                 float* dst, int64_t dst_t_stride, int64_t dst_n_stride,
                 int F, int Tc, void* stream);

int sgp_spmm_rbu_halo(const int32_t* grp_ptr, const int32_t* grp_rows, const int32_t* ucol,
                      const float* uval, int R, int n_groups,
                      const float* src, int64_t src_t_stride, int64_t src_n_stride,
                      const float* src2, int64_t src2_t_stride, int64_t src2_n_stride, int n_split,
                      float* dst, int64_t dst_t_stride, int64_t dst_n_stride,
                      int F, int Tc, void* stream);

/* Tensor-core hop (tcgen05, 3xTF32, fp32-accurate): rows grouped 64 at a time, union columns padded
 * to chunks of 32.  chunk_ptr [n_groups+1] (in chunks), grp_rows [n_groups, 64] (-1 = padding),
 * cols [total_chunks*32] source row ids, bimg [total_chunks][64*32] the chunk's operator values
 * (fp32) in the K-major SWIZZLE_128B shared-memory layout (built by sgp_b200/ops.py::tc_build; the
 * kernel splits them into tf32 hi / lo).  F in {128, 256, 512}.  *err_flag (device int) is set to 1 if
 * an internal barrier times out.  src2 / n_split as in sgp_spmm_halo.  checksum as in
 * sgp_reservoir_scan_tc (NULL = off).  Source row counts are unbounded: row * stride is formed in
 * 64 bits (row strides themselves must be < 2^32 bytes). */
int sgp_spmm_rbu_tc(const int32_t* chunk_ptr, const int32_t* grp_rows, const int32_t* cols,
                    const float* bimg, int n_groups,
                    const float* src, int64_t src_t_stride, int64_t src_n_stride,
                    const float* src2, int64_t src2_t_stride, int64_t src2_n_stride, int n_split,
                    float* dst, int64_t dst_t_stride, int64_t dst_n_stride,
                    int F, int Tc, int* err_flag, double* checksum /*nullable*/, void* stream);

/* Tensor-core hop with fp16x3 operands and 96-row groups (tcgen05 kind::f16; see csrc/spmm_tc16.cu):
 * same contract as sgp_spmm_rbu_tc.  grp_rows [n_groups, 96]; bimg [total_chunks][96*64] fp16 — per chunk
 * the [96 rows x 32 columns] slab of operator values times w_scale, split into hi | lo halves stored side
 * by side in 128-byte rows, K-major SWIZZLE_128B (built by sgp_b200/ops.py::tc16_build).  x_scale: a power
 * of two with x_scale * max|src| <= 2^14 (the caller must know a bound on the panel: 1 for tanh reservoir
 * states and their row-stochastic propagations); w_scale: the operator's own power-of-two scale. */
int sgp_spmm_rbu_tc16(const int32_t* chunk_ptr, const int32_t* grp_rows, const int32_t* cols,
                      const void* bimg, int n_groups,
                      const float* src, int64_t src_t_stride, int64_t src_n_stride,
                      const float* src2, int64_t src2_t_stride, int64_t src2_n_stride, int n_split,
                      float* dst, int64_t dst_t_stride, int64_t dst_n_stride,
                      int F, int Tc, float x_scale, float w_scale, int* err_flag,
                      double* checksum /*nullable*/, void* stream);

/* Process-wide limit on the persistent CTAs of sgp_spmm_rbu_tc / sgp_spmm_rbu_tc16 (default and maximum: one per SM, 148).
 * The row-sharded encoder lowers it on >= 4 GPUs so that a few SMs stay free for the halo push and
 * the barrier kernels, which cannot share an SM with a hop CTA (registers) and would otherwise wait
 * for the gap between two hop launches. */
int sgp_tc_set_cta_limit(int n_ctas);

/* HOST function (pointers are host memory, no stream): choose the R-row groups of the RBU format
 * from a CSR operator by a breadth-first, heaviest-neighbour-first greedy (group_rows.cu).
 * grp_rows must hold ceil(N/R)*R entries; unused slots of the last group are set to -1. */
int sgp_group_rows(const int32_t* rowptr, const int32_t* col, const float* val, int32_t N, int32_t R,
                   int32_t* grp_rows, int32_t* n_groups_out);

/* ---------------------------------------------------------------------------------------------
 * DynGESN layer update (lib/nn/reservoir/graph_reservoir.py:85-93), one time step, one layer:
 *     h' = (1 - alpha) h + alpha * act( x W_ih^T + b + prop ),   prop = S (h W_hh^T)
 * `prop` [N, H] is produced by the caller with sgp_reservoir_scan (identity, one step: h W_hh^T)
 * and the K2 SpMM; this fuses the input projection, bias, activation and leaky blend.  h_state
 * [N, H] contiguous in/out; out row n at out + n*out_n_stride (a feature block of the [T, N, L*H]
 * output); bias may be NULL.
 * ------------------------------------------------------------------------------------------- */
int sgp_gesn_update(const float* x, int64_t x_n_stride, int Fin, const float* w_ih /*[H,Fin]*/,
                    const float* bias /*[H] or NULL*/, const float* prop, int64_t prop_n_stride,
                    float alpha, float one_minus_alpha, int act, float* h_state,
                    float* out, int64_t out_n_stride, int N, int H, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Grouped 1x1 convolution (forward) — the first decoder layer over the encoder's (hop, layer)
 * feature blocks: nn.Conv1d(groups*Cin, groups*Cout, kernel_size=1, groups=groups) on 'b f n'
 * (lib/nn/models/sgp_model.py:41-52).  x row r at x + r*x_row_stride holds groups*Cin features,
 * weight [groups*Cout, Cin] (the Conv1d weight with its trailing kernel dim of 1 dropped), bias
 * [groups*Cout] or NULL, y row r at y + r*y_row_stride receives groups*Cout values.  Cout <= 256.
 * ------------------------------------------------------------------------------------------- */
int sgp_grouped_linear(const float* x, int64_t x_row_stride, const float* weight, const float* bias,
                       float* y, int64_t y_row_stride, int64_t rows, int groups, int Cin, int Cout,
                       void* stream);

/* HOST function: owner[i] in [0, parts) for every row — `parts` compact, equally sized patches of
 * the graph by recursive bisection along the patch diameter (group_rows.cu).  The row-sharded
 * encoder (no counterpart in the single-process reference) gives patch r to rank r. */
int sgp_partition_rows(const int32_t* rowptr, const int32_t* col, int32_t N, int32_t parts,
                       int32_t* owner /*[N]*/);

/* ---------------------------------------------------------------------------------------------
 * K4  global block: dst[t, n, :] = mean over nodes of src[t, :, :]
 * Replaces `torch.ones_like(x) * x.mean(-2, keepdim=True)`
 * (lib/nn/encoders/sgp_spatial_encoder.py:32-34).  `sums` is a [Tc, F] fp32 workspace; when
 * `precomputed_sums` != 0 it already holds the node SUMS over all shards (after an all-reduce) and
 * only the broadcast runs; N_total is the divisor.
 * ------------------------------------------------------------------------------------------- */
int sgp_node_sum(const float* src, int64_t src_t_stride, int64_t src_n_stride,
                 float* sums /*[Tc,F]*/, int N, int F, int Tc, void* stream);
int sgp_node_mean_broadcast(const float* sums /*[Tc,F]*/, int64_t N_total,
                            float* dst, int64_t dst_t_stride, int64_t dst_n_stride,
                            int N, int F, int Tc, void* stream);

/* Output sink for streamed benchmarking: acc[0] += sum(buf[0:count]) in fp64 (device scalar). */
int sgp_checksum(const float* buf, int64_t count, double* acc, void* stream);

/* The same over a strided [Tc, N, F] view (a feature block of the encoder's output buffer). */
int sgp_checksum_view(const float* src, int64_t src_t_stride, int64_t src_n_stride, int N, int F, int Tc,
                      double* acc, void* stream);

/* Halo push (row-sharded path): dst_addr[k] + t*dst_t_stride receives src[t, index[k], 0:F] for every
 * t < Tc.  dst_addr [n_index] is a DEVICE array of 64-bit addresses, normally slots of peer GPUs'
 * halo buffers mapped into this process: the pack and the NVLink transfer are one kernel.
 * F % 4 == 0, 16-byte aligned rows.  Ordering against the readers is the caller's business
 * (a cross-rank barrier before and after, sgp_b200/sharded.py). */
int sgp_push_rows(const float* src, int64_t src_t_stride, int64_t src_n_stride,
                  const int32_t* index, const int64_t* dst_addr, int n_index,
                  int64_t dst_t_stride, int F, int Tc, void* stream);

/* IID (t, n) sampler gather: dst[m, 0:F] = src[t_idx[m], n_idx[m], 0:F] for m in [0, M).
 * Replaces `tens[(step_index, None, None, node_index)]` and `tens[(hor_index, node_index[:, None], None)]`
 * of IIDDataset.sample (lib/datasets/iid_dataset.py:57-99) for a device-resident `tens`.  The index
 * arrays are DEVICE int64 (the reference's torch.randint dtype); out-of-range indices are the
 * caller's responsibility (the reference would raise an IndexError on the host). */
int sgp_gather_tn(const float* src, int64_t src_t_stride, int64_t src_n_stride, int T, int N, int F,
                  const int64_t* t_idx, const int64_t* n_idx, int64_t M,
                  float* dst, int64_t dst_m_stride, void* stream);

/* Gather rows: dst[t, i, :] = src[t, index[i], :]  (halo packing for the row-sharded path). */
int sgp_gather_rows(const float* src, int64_t src_t_stride, int64_t src_n_stride,
                    const int32_t* index, int n_index,
                    float* dst, int64_t dst_t_stride, int64_t dst_n_stride,
                    int F, int Tc, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* SGP_B200_H */
