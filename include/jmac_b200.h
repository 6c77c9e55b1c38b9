/*
 * jmac_b200.h -- C ABI of libjmac_b200.so: hand-written sm_100a CUDA kernels for JMAC's hot path.
 *
 * Conventions
 *   - plain C, no torch / CUDA types: device pointers are ordinary pointers, `stream` is a
 *     cudaStream_t passed as void* (NULL = legacy default stream).
 *   - every function returns 0 on success, a positive cudaError_t value on a CUDA failure, or a
 *     negative JMAC_E* code on an argument error.  Nothing throws, nothing synchronises the stream
 *     (except plan creation and the *_host convenience entry points, which say so).
 *   - all float tables are fp32 row-major; `d` (row width) must be a multiple of 4, 4 <= d <= 512,
 *     rows 16-byte aligned.  Index inputs that mirror the reference are int64 (torch.LongTensor).
 *   - the library never allocates through torch; scratch memory lives in the plan / is passed in.
 *
 * Each entry point cites the reference code (file:line under the reference tree) it replaces.
 */
#ifndef JMAC_B200_H
#define JMAC_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define JMAC_OK 0
#define JMAC_EINVAL (-1)   /* bad argument (null pointer, d not supported, sizes out of range) */
#define JMAC_ERANGE (-2)   /* an index in edge_index / edge_type is out of range              */
#define JMAC_ENOMEM (-3)

const char* jmac_version(void);
/* Human-readable text for the last non-zero return value on this thread. */
const char* jmac_last_error(void);
/* Number of kernels this library has launched in this process (for bench.py's gpu_launches). */
uint64_t jmac_launch_count(void);
/* Asynchronous index errors.  The scorer / pair-cosine / link-prediction kernels take index lists straight from the caller
 * (the reference raises IndexError / a device-side assert on a bad one).  They never read or write out of bounds: an index
 * outside its table is replaced by row 0 and a flag is raised in mapped host memory.  jmac_status_poll returns and clears
 * the flags WITHOUT synchronising (bit 0: triple scorer h/r/t, bit 1: pair cosine ia/ib, bit 2: link prediction e/r/t/filter,
 * bit 3: negative sampler: (h, r) out of range or fewer than k valid entities);
 * synchronise the stream first to be sure the kernels in question have run. */
int jmac_status_poll(void);
/* Withhold n_sms SMs from the persistent (one CTA per SM) kernels of this library, process-wide; returns the previous value.
 * For callers that overlap a communication kernel (NCCL) with them: a persistent CTA fills its SM's shared memory, so a
 * collective launched meanwhile could otherwise only start in the gaps between kernels. */
int jmac_set_sm_reserve(int n_sms);

/* ------------------------------------------------------------------------------------------
 * Peer memory for the multi-GPU exchange (one process per GPU on one node; no counterpart in the single-device reference).
 * jmac_peer_alloc: a cudaMalloc allocation other processes can map; jmac_ipc_get_handle / jmac_ipc_open / jmac_ipc_close: its
 * 64-byte CUDA IPC handle, the mapping in a peer process, and its release; jmac_memcpy_async: a device-to-device copy on `stream`
 * (local or peer-mapped pointers) executed by the copy engines -- no SM is involved, so the transfer really overlaps the
 * persistent compute kernels.  Ordering across processes is the caller's job (jmac_b200/peer.py uses tiny NCCL all-reduces).
 * ------------------------------------------------------------------------------------------ */
int jmac_peer_alloc(size_t bytes, void** out);
int jmac_peer_free(void* ptr);
int jmac_ipc_get_handle(void* ptr, unsigned char* handle64);
int jmac_ipc_open(const unsigned char* handle64, void** out);
int jmac_ipc_close(void* ptr);
int jmac_memcpy_async(void* dst, const void* src, size_t bytes, void* stream);
/* Cross-process ordering without an SM (stream memory operations): jmac_stream_write_u32 stores `value` at a LOCAL device word in
 * stream order; jmac_stream_wait_geq_u32 blocks `stream` until (int32)(*addr - value) >= 0 for a LOCAL device word (peers set it
 * with a 4-byte jmac_memcpy_async).  jmac_peer_sum: the local half of the peer reduce-scatter, dst[i] (+)= sum over g <
 * n_slices of src[g * slice_stride + i], i < n, slices added in ascending order (bit-reproducible). */
int jmac_stream_write_u32(void* addr, unsigned int value, void* stream);
int jmac_stream_wait_geq_u32(void* addr, unsigned int value, void* stream);
int jmac_peer_sum(float* dst, const float* src, int n_slices, int64_t slice_stride, int64_t n, int accumulate, void* stream);

/* ------------------------------------------------------------------------------------------
 * Graph plan: CSR-by-destination (+ by-source and by-relation orderings used by backward).
 * Replaces the per-call index work of RelationAwareLayer.compute_norm (src/jmac_model.py:99-109)
 * and the gathers of MessagePassing.propagate (modules/helper/message_passing.py:69-81).  Built
 * once per edge set; the reference reuses one edge set for thousands of steps (train.py:447-485).
 *   row = edge_index[0] (aggregation target i), col = edge_index[1] (message source j).
 *   n_rel = number of relation rows INCLUDING the self-loop row (= rel_emb.size(0) + 1).
 *   chunk = max edges one warp handles before a segment is split; 0 = automatic: 256 on large graphs, down to 32 on the
 *           reference's graph sizes, where a warp's fair share of the edges is a few dozen.
 * Synchronises `stream` (it needs data-dependent sizes); the plan is complete on return and may be
 * used from any stream.  Its arrays come from the device's default stream-ordered memory pool
 * (cudaMallocAsync; the pool keeps up to 8 GiB of freed blocks per device, so a rebuilt plan reuses the
 * memory of the destroyed one).  jmac_plan_destroy does not synchronise the device: every entry point
 * records an event behind its last kernel, the frees are ordered behind those events on a private stream
 * and the call waits for that stream only.  A plan owns the scratch for the partial results of
 * split segments: edge kernels of ONE plan must not run concurrently on different streams.
 * ------------------------------------------------------------------------------------------ */
typedef struct jmac_plan jmac_plan;

int jmac_plan_create(const int64_t* d_row, const int64_t* d_col, const int64_t* d_type, int64_t E,
                     int64_t N, int64_t n_rel, int chunk, void* stream, jmac_plan** out);
/* Destination-sharded variant (multi-GPU): row indexes the N_dst nodes this shard aggregates into,
 * col indexes a (gathered) source table of N_src rows.  P/agg/lse/G/dP are [N_dst,d]; Q/X and the
 * source-pass outputs dXe/dQ are [N_src,d]. */
int jmac_plan_create_bipartite(const int64_t* d_row, const int64_t* d_col, const int64_t* d_type,
                               int64_t E, int64_t N_dst, int64_t N_src, int64_t n_rel, int chunk,
                               void* stream, jmac_plan** out);
int jmac_plan_destroy(jmac_plan* plan);
/* info[0..7] = E, N, n_rel, chunk, #dst work items, #src work items, #rel work items, max in-degree */
int jmac_plan_info(const jmac_plan* plan, int64_t info[8]);
/* Copies (device -> caller's DEVICE buffers, any may be NULL): rowptr[N+1], csr_src[E],
 * csr_type[E], csr_eid[E] (position -> original edge id), sqrt_deg[N]. */
int jmac_plan_export(const jmac_plan* plan, int32_t* rowptr, int32_t* csr_src, int32_t* csr_type,
                     int32_t* csr_eid, float* sqrt_deg, void* stream);

/* ------------------------------------------------------------------------------------------
 * Edge phase of RelationAwareLayer.forward for comp_op='sub'
 * (src/jmac_model.py:49,56-89 + modules/helper/message_passing.py:4-29), in the factored form
 *   s_e   = a_att . LeakyReLU(P[i] + Q[j] - S[t]),  P = X W_top, Q = X W_bot, S = R' W_bot
 *   alpha = softmax over the in-edges of i;  agg[i] = sqrt(deg_i) * sum_e alpha_e (X[j] - R'[t])
 * (the division by edge_norm = deg^-1/2 of message_passing.py:25-26 is the sqrt(deg) factor).
 * Outputs: agg[N,d]; lse[N] = log-sum-exp of the logits of each destination (saved for backward).
 * ------------------------------------------------------------------------------------------ */
int jmac_edge_fwd(jmac_plan* plan, const float* P, const float* Q, const float* X, const float* S,
                  const float* Rp, const float* a_att, float leaky_slope, int d, float* agg,
                  float* lse, void* stream);

/* Same, with two optional fusions:
 *  - Xself/T (both NULL or both set): the self-loop propagate of src/jmac_model.py:50 is folded into
 *    the epilogue, T[i] = agg[i] + Xself[i] - Rp[n_rel-1] (the operand of the message GEMM), Xself
 *    being the [N_dst,d] table of the destinations' own embeddings;
 *  - erec (NULL when no gradient is needed): the per-edge record buffer of
 *    E * jmac_edge_record_bytes(d) bytes; forward stores the attention logit s_e and the sign mask of
 *    z_e = P[i]+Q[j]-S[t] (1 bit per feature) of every edge, which is all backward needs of P, Q, S. */
int jmac_edge_fwd_fused(jmac_plan* plan, const float* P, const float* Q, const float* X,
                        const float* S, const float* Rp, const float* a_att, float leaky_slope, int d,
                        float* agg, float* lse, const float* Xself, float* T, void* erec,
                        void* stream);

/* Bytes per edge of the record buffer `erec` shared by forward and the backward passes
 * (sign mask 32/64 B + {alpha*sqrt(deg), ds_e} 8 B + logit 4 B).  The buffer must be 32-byte aligned. */
size_t jmac_edge_record_bytes(int d);

/* Backward of src/jmac_model.py:56-89 + message_passing.py:24-28 in three atomics-free segmented
 * passes, each gathering ONE node row per edge.  G = dL/d agg ([N_dst,d]).  Every pass ACCUMULATES its
 * term of da_att[d] (caller zeroes it once): LeakyReLU is positively homogeneous, so
 * d a = sum_i P[i](.)mu_i + sum_j Q[j](.)nu_j - sum_t S[t](.)omega_t with the per-segment sums of
 * ds_e * LeakyReLU'(z_e) that the passes form anyway.
 * Destination pass: reads the forward's logits and masks from erec, gathers X[j]; writes dP[N_dst,d]
 * and the {alpha*sqrt(deg), ds_e} part of erec.  g_colsum[d] (may be NULL; caller zeroes): ACCUMULATES the column sums of G
 * over the destinations -- the gradient of the self-loop relation row of the fused forward (T = agg + X - R'[loop]), picked up
 * while every G row passes through anyway instead of by a separate reduction over the [N,d] table. */
int jmac_edge_bwd_dst(jmac_plan* plan, const float* P, const float* X, const float* Rp,
                      const float* a_att, float leaky_slope, int d, const float* agg, const float* lse,
                      const float* G, float* dP, void* erec, float* da_att, float* g_colsum,
                      void* stream);
/* Source pass (by-source ordering): dXe[N_src,d] = sum over out-edges of alpha*sqrt(deg)*G[i] (the
 * x_j gather's gradient, replaces index_put_(accumulate)); dQ[N_src,d] = a_att (.) nu_j. */
int jmac_edge_bwd_src(jmac_plan* plan, const float* G, const void* erec, const float* a_att,
                      float leaky_slope, int d, const float* Q, float* dXe, float* dQ, float* da_att,
                      void* stream);
/* Relation pass (by-relation ordering): dRp[n_rel,d] = -sum alpha*sqrt(deg)*G[i],
 * dS[n_rel,d] = -a_att (.) omega_t.  Tables are overwritten (zeroed inside). */
int jmac_edge_bwd_rel(jmac_plan* plan, const float* G, const void* erec, const float* a_att,
                      float leaky_slope, int d, const float* S, float* dRp, float* dS, float* da_att,
                      void* stream);

/* ------------------------------------------------------------------------------------------
 * BatchNorm1d over the N nodes + tanh (src/jmac_model.py:52), split so that a multi-GPU caller can
 * all-reduce the per-feature sums between the two halves.
 * ------------------------------------------------------------------------------------------ */
/* sums[2*d] (double, device): ACCUMULATES sum_x[d], sum_x2[d] of pre[N,d] (caller zeroes). */
int jmac_bn_stats(const float* pre, int64_t N, int d, double* sums, void* stream);
/* training != 0: uses sums over n_total rows, writes save_mean[d], save_rstd[d], updates
 * running_mean/var (unbiased variance, momentum) when they are non-NULL.
 * training == 0: uses running_mean/var.  out = tanh(gamma * xhat + beta). */
int jmac_bn_tanh_apply(const float* pre, int64_t N, int d, const double* sums, int64_t n_total,
                       const float* gamma, const float* beta, float* running_mean,
                       float* running_var, float momentum, float eps, int training, float* out,
                       float* save_mean, float* save_rstd, void* stream);
/* bsums[2*d] (double): ACCUMULATES sum dy[d], sum dy*xhat[d] with dy = gout * (1 - out^2). */
int jmac_bn_tanh_bwd_stats(const float* gout, const float* out, const float* pre, int64_t N, int d,
                           const float* mean, const float* rstd, double* bsums, void* stream);
/* gpre = gamma*rstd*(dy - sum_dy/n - xhat*sum_dyxhat/n) (training) or gamma*rstd*dy (eval);
 * dgamma[d] = sum dy*xhat, dbeta[d] = sum dy (written from bsums when non-NULL). */
int jmac_bn_tanh_bwd_apply(const float* gout, const float* out, const float* pre, int64_t N, int d,
                           const float* mean, const float* rstd, const float* gamma,
                           const double* bsums, int64_t n_total, int training, float* gpre,
                           float* dgamma, float* dbeta, void* stream);

/* ------------------------------------------------------------------------------------------
 * Row-wise L2 normalisation y = x / max(|x|_2, eps) (F.normalize of src/jmac_model.py:183,192,230-231) and its backward,
 * one pass each.  inv[N] (may be NULL when no gradient is needed) keeps 1 / max(|x|, eps), negative where the clamp was active.
 * ------------------------------------------------------------------------------------------ */
int jmac_row_normalize_fwd(const float* x, int64_t N, int d, float eps, float* y, float* inv, void* stream);
int jmac_row_normalize_bwd(const float* y, const float* inv, const float* g, int64_t N, int d, float* dx,
                           void* stream);

/* ------------------------------------------------------------------------------------------
 * Node-level dense product with fp32 accuracy on the tensor cores (3xTF32 split, tcgen05):
 *   C[M,N] = A[M,K] . B[N,K]^T,  B given pre-split: B_hi = B with the low 13 mantissa bits cleared,
 *   B_lo = B - B_hi.  N % 32 == 0, N <= 512 (computed as column parts of <= 256), K % 16 == 0; lda/ldb/ldc in elements (multiples of 4).
 * Replaces the fp32 library GEMMs of the hoisted per-edge products of src/jmac_model.py:75-76,88
 * (P = X W_top, Q = X W_bot, pre = T W_g) and their input-gradient transposes.
 * ------------------------------------------------------------------------------------------ */
int jmac_gemm_tf32x3(const float* A, int64_t lda, const float* B_hi, const float* B_lo, int64_t ldb,
                     float* C, int64_t ldc, int64_t M, int N, int K, void* stream);
/* The pre-split of the small operand in one pass: W[rows, cols] (leading dimension ld) -> hi = RN-to-TF32(w), lo = w - hi, both
 * contiguous, laid out as W (transpose == 0) or as W^T [cols, rows] (transpose != 0: Y = X W needs B = W^T in [N,K] form). */
int jmac_split_tf32(const float* W, int64_t rows, int64_t cols, int64_t ld, int transpose, float* hi,
                    float* lo, void* stream);
/* Same product with up to two [M,N] addends folded into the epilogue, C = A.B^T + D1 (+ D2); D1/D2
 * may alias C.  Fuses the gradient accumulation of the layer input (autograd's adds of
 * dP W_top^T, dQ W_bot^T, the edge pass's dX and the self-loop term, src/jmac_model.py:49-52,75-76). */
int jmac_gemm_tf32x3_add(const float* A, int64_t lda, const float* B_hi, const float* B_lo,
                         int64_t ldb, float* C, int64_t ldc, int64_t M, int N, int K, const float* D1,
                         const float* D2, int64_t ldd, void* stream);
/* Weight-gradient product D[KO,N] = X[M,KO]^T . G[M,N] (reduction over the M node rows), same
 * 3xTF32 scheme with both operands split in shared memory; KO % 32 == 0, KO <= 256, N % 32 == 0, N <= 256 (wider tables: one call
 * per column block of X and of G, jmac_b200.functional.gemm_tn).
 * Deterministic: per-CTA partials in `workspace` (jmac_gemm_tn_workspace_bytes) summed in order. */
size_t jmac_gemm_tn_workspace_bytes(int KO, int N);
int jmac_gemm_tn_tf32x3(const float* X, int64_t ldx, const float* G, int64_t ldg, float* D, int64_t M,
                        int KO, int N, float* workspace, void* stream);

/* ------------------------------------------------------------------------------------------
 * Completion scorer: TransE-L1 margin loss of JMAC.completion_loss (src/jmac_model.py:345-374)
 * on the reference's exact batch layout (train.py:347-352): M = (1+K)*B triples, positives are the
 * first B, the negative at flat position B + k*B + b is compared with positive b.
 *   score[l*M + m] = || ent_l[h_m] + rel_l[r_m] - ent_l[t_m] ||_1          (optional output, may be NULL)
 *   loss[l]        = mean_{b,k} max(score_b - score_{B+kB+b}, -margin) + margin
 * n_layers <= 4 tables are scored in one pass over the indices.  h and r need NOT be tiled (any index lists of length M
 * are accepted, as in the reference); the tiled layout is the fast case.
 * `record` (NULL when no gradient is needed; jmac_transe_record_bytes bytes, 256-byte aligned) receives what backward
 * needs: per triple d loss/d score up to the scale and the sign pattern of h + r - t (2 bits per feature), per positive the
 * dense gradient row of its leading (h, r) run.  Backward never gathers an embedding row again.
 * ------------------------------------------------------------------------------------------ */
size_t jmac_transe_record_bytes(int64_t M, int64_t B, int n_layers, int d);
int jmac_transe_margin_fwd(const float* const* ent, const float* const* rel, int n_layers, int d,
                           const int64_t* h, const int64_t* r, const int64_t* t, int64_t M,
                           int64_t B, int64_t n_ent, int64_t n_rel, float margin, float* score,
                           float* loss, void* record, void* stream);
/* Atomics-free, bit-reproducible backward: WRITES the rows of d_ent[l][n_ent,d], d_rel[l][n_rel,d] that occur in the batch
 * with gloss[l] * dloss[l]/d(row) (every such row exactly once); rows that do not occur are not touched (the caller provides
 * zeroed tables).  The occurrences of every row are radix-sorted once per call (all layers share the indices) and reduced in
 * a fixed order.  workspace: jmac_transe_bwd_workspace_bytes bytes, 256-byte aligned.  2*M + B < 2^30. */
size_t jmac_transe_bwd_workspace_bytes(int64_t M, int64_t B, int n_layers, int d);
int jmac_transe_margin_bwd(int n_layers, int d, const int64_t* h, const int64_t* r, const int64_t* t,
                           int64_t M, int64_t B, int64_t n_ent, int64_t n_rel, const void* record,
                           const float* gloss, float* const* d_ent, float* const* d_rel,
                           void* workspace, size_t workspace_bytes, void* stream);

/* ------------------------------------------------------------------------------------------
 * Filtered negative sampling: TrainDataset.get_neg_sample (modules/load/data_loader.py:36-47) for a whole batch on the
 * device.  keys[n_keys]: the SORTED, unique int64 keys (h*n_rel + r)*n_ent + t of all known triples.  For every row b,
 * out[b, 0..k) receives k DISTINCT entities drawn uniformly from those e with (h[b], r[b], e) not a known triple
 * (sequential rejection of a Philox-4x32-10 stream keyed by (seed, b): the reference's distribution, not its stream).
 * A row that cannot be filled (fewer than k valid entities, index out of range) raises status bit 3 (jmac_status_poll).
 * ------------------------------------------------------------------------------------------ */
int jmac_sample_negatives(const int64_t* keys, int64_t n_keys, const int64_t* h, const int64_t* r,
                          int64_t B, int64_t n_ent, int64_t n_rel, int k, uint64_t seed, int64_t* out,
                          void* stream);

/* ------------------------------------------------------------------------------------------
 * Pair cosine: cos[m] = <A[ia_m], B[ib_m]> / (max(|A[ia_m]|,1e-12) max(|B[ib_m]|,1e-12)), the
 * F.normalize + row-dot of JMAC.alignment_loss_simple / alignment_loss (src/jmac_model.py:243-248,
 * 268-287).  Backward ACCUMULATES into dA, dB (caller zeroes).
 * ------------------------------------------------------------------------------------------ */
int jmac_pair_cos_fwd(const float* A, const float* B, int d, const int64_t* ia, const int64_t* ib,
                      int64_t M, int64_t nA, int64_t nB, float* cosv, void* stream);
int jmac_pair_cos_bwd(const float* A, const float* B, int d, const int64_t* ia, const int64_t* ib,
                      int64_t M, int64_t nA, int64_t nB, const float* gcos, float* dA, float* dB,
                      void* stream);

/* ------------------------------------------------------------------------------------------
 * Link prediction: JMAC.forward_linkpred (src/jmac_model.py:301-313) and the ranking loop of
 * CompletionEvaluator.test (src/validate.py:50-70).
 *   dist[b,n] = sum_l || ent_l[e_b] + sign*rel_l[r_b] - ent_l[n] ||_1     (sign=-1 for pred_head)
 * jmac_l1_dist writes the [B,N] matrix (API parity); jmac_l1_rank never materialises it: it
 * returns rank_b = 1 + #{n : -dist[b,n] > -dist[b,t_b], n not filtered} where the filter of query b
 * is filt_idx[filt_ptr[b] .. filt_ptr[b+1]) (NULL = raw setting); the target itself never counts.
 * ------------------------------------------------------------------------------------------ */
int jmac_l1_dist(const float* const* ent, const float* const* rel, int n_layers, int d,
                 const int64_t* e_idx, const int64_t* r_idx, int64_t B, int64_t N, int64_t n_rel,
                 int pred_head, float* dist, void* stream);
int jmac_l1_rank(const float* const* ent, const float* const* rel, int n_layers, int d,
                 const int64_t* e_idx, const int64_t* r_idx, const int64_t* t_idx, int64_t B,
                 int64_t N, int64_t n_rel, int pred_head, const int64_t* filt_ptr,
                 const int64_t* filt_idx, int32_t* rank, void* stream);

/* ------------------------------------------------------------------------------------------
 * Alignment scorer: similarity GEMM on the 5th-gen tensor cores with a streaming epilogue
 * (modules/finding/similarity.py:13-84 sim / csls_sim / calculate_nearest_k,
 *  modules/finding/alignment.py:87-112 calculate_rank, modules/utils/util.py:31-54 get_neg).
 * The n1 x n2 matrix is never materialised.
 * ------------------------------------------------------------------------------------------ */
/* Two operand precisions.
 *   bf16 (jmac_sim_topk, jmac_sim_softmax_stats): rows rounded to bf16, one kind::f16 MMA per K step; the mode for
 *     the 1M x 1M problem of BASELINE.json configs[4].  dpad multiple of 64, <= 320.
 *   fp32-accurate 3xTF32 (the *_f32 entry points): every table is passed as two fp32 tables hi + lo produced by
 *     jmac_to_tf32_rows; three kind::tf32 MMAs per K step; |S - exact| ~ 1e-6 sum|a||b|, i.e. fp32-GEMM accuracy.  This is
 *     what the reference computes (f64 cdist -> f32, similarity.py:44-46; fp32 mm, util.py:52; train.py:239-257) and the
 *     default of the Python host for test / valid / get_neg / compute_alignment_quality.  dpad multiple of 32, <= 4096.
 * Small problems (< 4 row blocks of 256 per SM) are cut into one equal run of 256x128 tiles per SM so that the sizes the
 * reference evaluates (2.6k-30k rows) use every SM; the per-piece partial results live in `workspace`
 * (jmac_sim_workspace_bytes; 16-byte aligned; NULL or too small = unsplit launch, at most one CTA per 256 rows) and
 * are merged in ascending column order: top-k, counts and the gold entry are bit-identical to the unsplit sweep. */
int jmac_sm_count(void);
size_t jmac_sim_workspace_bytes(int64_t n1, int64_t n2, int k);
/* CTAs the similarity kernel launches for an n1 x n2 problem with (split != 0) / without a workspace. */
int jmac_sim_grid(int64_t n1, int64_t n2, int split);

/* fp32 rows src[idx[r]] (idx NULL = identity), optionally L2-normalised (cosine metric), rounded to
 * bf16 and zero-padded to dpad (multiple of 64) columns: dst[n_rows, dpad]. */
int jmac_to_bf16_rows(const float* src, int d, const int64_t* idx, int64_t n_rows, int normalize,
                      uint16_t* dst, int dpad, void* stream);
/* Same gather / normalisation (norm accumulated in fp64), split for the 3xTF32 product:
 * hi = RN-to-TF32(v), lo = RN-to-TF32(v - hi), both [n_rows, dpad] fp32, dpad a multiple of 32. */
int jmac_to_tf32_rows(const float* src, int d, const int64_t* idx, int64_t n_rows, int normalize,
                      float* hi, float* lo, int dpad, void* stream);
/* S = A . B^T (inputs [n1,dpad], [n2,dpad], fp32 accumulation); per row i of A, over all j:
 *   v_ij = (alpha * S_ij - r1[i]) - r2[j]            (r1, r2 may be NULL = 0; r2 must be readable up to
 *                                                      the next multiple of 128 entries, 16-byte aligned)
 *   topv/topi[n1,k] : the k largest v_ij, descending, ties -> lower j     (k = 0: skipped, k <= 32)
 *   gold_cnt[i]     : #{ j : v_ij > gold_val[i] }                          (NULL: skipped)
 *   diag[i]         : raw S at column gold_col[i]                           (NULL: skipped) */
int jmac_sim_topk(const uint16_t* A, const uint16_t* B, int64_t n1, int64_t n2, int dpad, int k,
                  float alpha, const float* r1, const float* r2, const int32_t* gold_col,
                  const float* gold_val, int32_t* gold_cnt, float* diag, float* topv,
                  int32_t* topi, void* workspace, size_t workspace_bytes, void* stream);
int jmac_sim_topk_f32(const float* A_hi, const float* A_lo, const float* B_hi, const float* B_lo,
                      int64_t n1, int64_t n2, int dpad, int k, float alpha, const float* r1,
                      const float* r2, const int32_t* gold_col, const float* gold_val,
                      int32_t* gold_cnt, float* diag, float* topv, int32_t* topi, void* workspace,
                      size_t workspace_bytes, void* stream);
/* Softmax statistics of x_ij = alpha * S_ij per row i, never materialising the matrix
 * (compute_alignment_quality, train.py:231-259: softmax(simi * 20), its entropy, its row max/argmax).
 * Rows with row_mask[i] != 0 and columns with col_mask[j] != 0 take x = masked_value instead
 * (train.py:254-255 sets them to -1, i.e. masked_value = -alpha).  col_mask follows the r2 rules
 * (padded to a multiple of 128, 16-byte aligned); either mask may be NULL.  Outputs per row:
 *   out_max = max_j x, out_arg = first argmax, out_Z = sum_j e^(x - max), out_W = sum_j e^(x-max) (x-max)
 * so that  max probability = 1/Z  and  entropy = log Z - W/Z. */
int jmac_sim_softmax_stats(const uint16_t* A, const uint16_t* B, int64_t n1, int64_t n2, int dpad,
                           float alpha, const float* row_mask, const float* col_mask,
                           float masked_value, float* out_max, int32_t* out_arg, float* out_Z,
                           float* out_W, void* workspace, size_t workspace_bytes, void* stream);
int jmac_sim_softmax_stats_f32(const float* A_hi, const float* A_lo, const float* B_hi,
                               const float* B_lo, int64_t n1, int64_t n2, int dpad, float alpha,
                               const float* row_mask, const float* col_mask, float masked_value,
                               float* out_max, int32_t* out_arg, float* out_Z, float* out_W,
                               void* workspace, size_t workspace_bytes, void* stream);
/* out[i] = (alpha * diag[i] - r1[i]) - r2[gold_col[i]], with the exact operation order of the
 * epilogue above, so that the gold entry never counts as greater than itself. */
int jmac_csls_gold(float alpha, const float* diag, const float* r1, const float* r2,
                   const int32_t* gold_col, int64_t n, float* out, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* JMAC_B200_H */
