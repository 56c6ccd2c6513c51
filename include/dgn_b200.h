/*
 * dgn_b200.h - C ABI of the B200-native DGN directional-aggregation engine.
 *
 * The reference (Saro00/DGN) is pure Python and has no FFI of its own.  The entry points
 * below are what a binding for its hot path has to call; each one names the reference code
 * it replaces (paths relative to the reference root, rb/ = realworld_benchmark/):
 *
 *   dgn_agg_forward   - rb/nets/dgn_layer.py:86-98 (reduce_func: every AGGREGATORS entry of
 *                       rb/nets/aggregators.py:8-93 followed by every SCALERS entry of
 *                       rb/nets/scalers.py:7-21), the message construction of
 *                       rb/nets/dgn_layer.py:75-84 / :154-159, and the degree-bucketed
 *                       DGLGraph.update_all of DGL 0.4.2 called at rb/nets/dgn_layer.py:115,186,264.
 *   dgn_agg_backward  - what torch.autograd derives for the same code.
 *   dgn_build_csr     - the edge grouping DGL 0.4.2 does inside update_all (degree bucketing)
 *                       and dgl.batch (rb/data/molecules.py:229).
 *   dgn_field_build   - the eigenvector weights of every edge, rb/nets/aggregators.py:35-71 (the
 *                       |eig_s - eig_d| / sum ... expressions every directional aggregator of every
 *                       layer re-evaluates on the mailbox), and the copy of ndata['eig'] onto the edges in
 *                       rb/nets/dgn_layer.py:75-84: evaluated ONCE per batch and shared by all layers.
 *   dgn_norm_*        - rb/nets/dgn_layer.py:122-130 (graph norm, BatchNorm1d, ReLU, residual).
 *   dgn_readout_*     - dgl.{mean,sum,max}_nodes at rb/nets/molecules_graph_regression/dgn_net.py:71-86.
 *   dgn_gemm_tf32x3   - the Linear layers of pretrans / posttrans (FCLayer, rb/nets/layers.py:76-100) in fp32
 *                       accuracy on the tcgen05 tensor cores.
 *   dgn_head_*        - the graph-level prediction head MLPReadout (rb/nets/mlp_readout_layer.py:11-30) in one launch
 *                       per direction; dgn_l1_loss_* - nn.L1Loss of rb/nets/molecules_graph_regression/dgn_net.py:90-92.
 *   dgn_embedding_backward, dgn_adam_step - the two remaining per-step pieces of the training loop that sit
 *                       between launches of the path (rb/nets/molecules_graph_regression/dgn_net.py:58,
 *                       rb/main_molecules.py:82).
 *
 * Conventions: plain pointers and sizes only, no ownership transfer, no allocation, no
 * exceptions.  Every pointer in DgnGraph / DgnAggIO / DgnAggGrad is DEVICE memory unless the
 * function name ends in _host.  Matrices are row-major fp32 with an explicit leading
 * dimension counted in floats.  `stream` is a cudaStream_t passed as void* (0 = legacy
 * default stream).  All launches are stream-ordered, re-entrant and CUDA-graph capturable.
 * Return value: 0 on success, a negative DgnStatus otherwise (see dgn_status_string).
 */
#ifndef DGN_B200_H_
#define DGN_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DGN_ABI_VERSION 13
#define DGN_MAX_AGG 32      /* aggregators per layer (the reference registry has 24)        */
#define DGN_MAX_SCALERS 4   /* scalers per layer (the reference registry has 3)             */
#define DGN_MAX_SLOTS 8     /* distinct eigen-weighted feature sums one launch can carry    */
#define DGN_EPS 1e-8f       /* rb/nets/aggregators.py:5                                     */
#define DGN_NORM_WS_FLOATS(C) (640 * (C)) /* fp32 workspace of dgn_norm_* for C columns         */

typedef enum {
  DGN_OK = 0,
  DGN_ERR_INVALID = -1,     /* bad argument (null pointer, negative size, unknown enum)     */
  DGN_ERR_UNSUPPORTED = -2, /* valid request the kernels cannot serve (too many slots ...)  */
  DGN_ERR_ALIGNMENT = -3,   /* an operand that must be 16 B aligned is not                  */
  DGN_ERR_CUDA = -4         /* a CUDA runtime call failed; see dgn_last_cuda_error          */
} DgnStatus;

/* aggregator kinds: rb/nets/aggregators.py */
typedef enum {
  DGN_AGG_MEAN = 0,            /* :8   */
  DGN_AGG_SUM = 1,             /* :31  */
  DGN_AGG_MAX = 2,             /* :12  */
  DGN_AGG_MIN = 3,             /* :16  */
  DGN_AGG_STD = 4,             /* :20  */
  DGN_AGG_VAR = 5,             /* :24  */
  DGN_AGG_DIR_AV = 6,          /* :35  dirK-av (alias dirK-smooth)  */
  DGN_AGG_DIR_DX = 7,          /* :48  dirK-dx ("dx_abs")           */
  DGN_AGG_DIR_DX_NO_ABS = 8,   /* :55  */
  DGN_AGG_DIR_DX_BALANCED = 9, /* :62  */
  DGN_AGG_DIR_SOFTMAX = 10     /* :42  dirK-0.1 / dirK-neg-0.1, alpha in agg_alpha */
} DgnAggKind;

/* scaler kinds: rb/nets/scalers.py */
typedef enum {
  DGN_SCALE_IDENTITY = 0,      /* :7   */
  DGN_SCALE_AMPLIFICATION = 1, /* :11  h * log(D+1) / avg_log */
  DGN_SCALE_ATTENUATION = 2    /* :16  h * avg_log / log(D+1) */
} DgnScalerKind;

/* how the message m_uv of edge u->v is formed */
typedef enum {
  DGN_MSG_SOURCE = 0, /* m = X[u]                      DGNLayerSimple, rb/nets/dgn_layer.py:154-155        */
  DGN_MSG_AFFINE = 1, /* m = X[u] + Q[v] (+ R[eid])    1-layer pretrans split per node, :75-80             */
  DGN_MSG_DENSE = 2   /* m = R[eid]                    materialised messages (pretrans_layers > 1)         */
} DgnMsgMode;

/* Batched graph in destination-major CSR ("in-edge slots") plus its by-source transpose. */
typedef struct {
  int32_t n_nodes;
  int32_t n_edges;
  const int32_t* in_ptr;   /* [n_nodes+1] slot range of every destination node                         */
  const int32_t* in_src;   /* [n_edges]   source node of every slot; slots of one node keep edge-id order */
  const int32_t* in_eid;   /* [n_edges]   original edge id of every slot (NULL = identity)              */
  const int32_t* out_ptr;  /* [n_nodes+1] range of every source node in out_slot (backward only)        */
  const int32_t* out_slot; /* [n_edges]   in-edge slot of every out-edge, grouped by source             */
  const float* log_deg;    /* [n_nodes]   (float)log(in_degree + 1), the scalers' per-node factor       */
  /* optional (NULL / 0 = unknown): the graphs of the batch as contiguous node ranges.  With them, high-degree batches
   * run one CTA per graph with the graph's source rows staged in shared memory (the batched adjacency of dgl.batch,
   * rb/data/molecules.py:229, is block diagonal: every in-edge of a node comes from its own graph). */
  const int32_t* graph_ptr;/* [n_graphs+1] first node of every graph, graph_ptr[n_graphs] = number of real nodes */
  int32_t n_graphs;
  int32_t max_graph_nodes; /* an upper bound of the largest graph's node count (sizes the shared-memory tile)   */
} DgnGraph;

/* Eigen-field of one batch: the normalised eigenvector weight w_s(u->v) of every in-edge for every distinct
 * (eigenvector column, weight kind) pair ("slot") the spec's directional aggregators need, laid out in groups of
 * DGN_FIELD_GROUP in-edge slots for 128-bit loads.  Group g is (1 + n_slots) * 4 words:
 *   int32 src[4] (source node of the slot, -1 = padding), then float w_s[4] for s = 0 .. n_slots-1.
 * Node v owns group v (its first 4 in-edges) and the overflow groups N + ovf_ptr[v] .. N + ovf_ptr[v+1]-1.
 * Slots are numbered in order of first use by the spec's aggregator list: dirK-av -> |d|/(sum|d|+eps);
 * dirK-dx and dirK-dx-no-abs share d/(sum|d|+eps); dirK-dx-balanced; dirK-(neg-)0.1 -> softmax(alpha |d|).
 * All memory is caller-owned device memory; `groups` must be 16 B aligned.  Built by dgn_field_build, consumed
 * by dgn_agg_forward / dgn_agg_backward through DgnAggIO.field. */
#define DGN_FIELD_GROUP 4
typedef struct {
  int32_t n_groups;        /* capacity of `groups`: >= n_nodes + ovf_ptr[n_nodes]                              */
  int32_t n_slots;         /* dgn_field_slots(spec)                                                            */
  const int32_t* ovf_ptr;  /* [n_nodes+1] from dgn_build_groups_host                                           */
  float* groups;           /* [n_groups][1 + n_slots][4]                                                       */
  float* wsum;             /* [n_nodes][4*ceil(n_slots/4)]: sum of w_s over the in-edges of every node, 16 B aligned
                              (NULL if n_slots == 0)                                                           */
} DgnField;

/* Which aggregators / scalers to compute: the AGGREGATORS / SCALERS names resolved to op-codes. */
typedef struct {
  int32_t n_feat;         /* F: feature columns aggregated (all towers together)                        */
  int32_t group_feat;     /* columns per tower (== n_feat when there is a single tower)                 */
  int32_t n_eig;          /* K: columns of eig                                                          */
  int32_t n_agg;          /* A                                                                          */
  int32_t n_scalers;      /* S; like rb/nets/dgn_layer.py:95 the scalers are applied only when S > 1    */
  uint8_t agg_kind[DGN_MAX_AGG];   /* DgnAggKind                                                        */
  uint8_t agg_eig[DGN_MAX_AGG];    /* eigenvector column for directional kinds                          */
  float agg_alpha[DGN_MAX_AGG];    /* softmax temperature for DGN_AGG_DIR_SOFTMAX                       */
  uint8_t scaler_kind[DGN_MAX_SCALERS];
  float avg_log;          /* avg_d["log"] = mean(log(D+1)) over the training set                        */
} DgnAggSpec;

/* Operands of one aggregation.  Output row v, tower t, scaler s, aggregator a, column c lives at
 *   out[v*ld_out + t*out_group_stride + (s*A + a)*group_feat + c]
 * which for one tower is the reference's [N, S*A*F] layout (scaler-major, then aggregator). */
typedef struct {
  int32_t msg_mode;       /* DgnMsgMode                                                                 */
  const float* x;         /* [N,F] source term: X (SOURCE) or P = h W_src^T (AFFINE); unused for DENSE   */
  int32_t ld_x;
  const float* q;         /* [N,F] destination term (AFFINE only)                                        */
  int32_t ld_q;
  const float* q_bias;    /* [F] optional bias added to every message (AFFINE only): m = X[u] + Q[v] + q_bias */
  const float* r;         /* [E,F] per-edge term in EDGE-ID order: optional for AFFINE, the message for DENSE */
  int32_t ld_r;
  const float* h_in;      /* [N,F] destination's own features (the dx aggregators subtract W*h_in)       */
  int32_t ld_h;
  const float* eig;       /* [N,K]                                                                       */
  int32_t ld_eig;
  float* out;             /* see layout above                                                            */
  int32_t ld_out;
  int32_t out_group_stride;
  float* h_copy;          /* optional: h_in rows copied here (tower t at h_copy[v*ld_hcopy + t*hcopy_group_stride + c]),
                             fuses the torch.cat([h, agg]) of rb/nets/dgn_layer.py:116                    */
  int32_t ld_hcopy;
  int32_t hcopy_group_stride;
  const DgnField* field;  /* optional (host pointer to the struct): eigen-field of this batch built with a spec that
                             has the same directional aggregators.  With it the kernels read precomputed weights
                             (the fast path); NULL = weights are derived from eig inside every launch.          */
} DgnAggIO;

/* Gradients for dgn_agg_backward.  Any output pointer may be NULL when that gradient is not needed. */
typedef struct {
  const float* g_out;     /* gradient of `out`, same layout / ld_out / out_group_stride as the forward  */
  const float* g_hcopy;   /* gradient of `h_copy` (NULL if h_copy was not used), same layout            */
  float* d_x;             /* [N,F] gradient of io.x (scatter over out-edges; + d_h_in when fold_h_in)    */
  int32_t ld_dx;
  float* d_q;             /* [N,F] gradient of io.q                                                      */
  int32_t ld_dq;
  float* d_r;             /* [E,F] gradient of io.r in edge-id order                                     */
  int32_t ld_dr;
  float* d_h_in;          /* [N,F] gradient of io.h_in (incl. g_hcopy and d_h_addend)                    */
  int32_t ld_dh;
  const float* d_h_addend;/* [N,F] optional: added into d_h_in (e.g. the residual branch's gradient)     */
  int32_t ld_dha;
  float* edge_ws;         /* [E,F] workspace (slot order) for the deterministic source-side reduction;
                             required when d_x != NULL.  With d_x == NULL the per-edge message gradients are
                             only left here (the caller reduces them, see dgn_pair_gather_backward)      */
  int32_t fold_h_in;      /* 1: d_x += d_h_in (SOURCE mode where x and h_in are the same tensor)         */
} DgnAggGrad;

int dgn_abi_version(void);
const char* dgn_status_string(int status);
const char* dgn_last_cuda_error(void);

/* out[N, S*A*F] = scalers(aggregators(messages)) for every destination node.  Replaces
 * reduce_func + update_all + message construction (see file header). */
int dgn_agg_forward(const DgnGraph* g, const DgnAggSpec* spec, const DgnAggIO* io, void* stream);

/* Gradients of dgn_agg_forward w.r.t. x, q, r and h_in; recomputes the per-node statistics
 * instead of saving them.  Deterministic (no atomics). */
int dgn_agg_backward(const DgnGraph* g, const DgnAggSpec* spec, const DgnAggIO* io, const DgnAggGrad* grad,
                     void* stream);

/* HOST: group edges by destination (stable, so mailbox order = edge-id order as in DGL 0.4.2
 * degree bucketing) and build the by-source transpose.  All arrays are host memory sized as in
 * DgnGraph; log_deg may be NULL. */
int dgn_build_csr_host(int32_t n_nodes, int32_t n_edges, const int32_t* src, const int32_t* dst,
                       int32_t* in_ptr, int32_t* in_src, int32_t* in_eid, int32_t* out_ptr, int32_t* out_slot,
                       float* log_deg);

/* Number of eigen-field slots of a spec (0 .. DGN_MAX_SLOTS), negative DgnStatus on error. */
int dgn_field_slots(const DgnAggSpec* spec);

/* HOST: overflow-group offsets of the eigen-field layout from the in-edge ranges: node v gets
 * max(0, ceil((D_v - 4) / 4)) overflow groups.  Returns their total (>= 0) or a negative DgnStatus. */
int dgn_build_groups_host(int32_t n_nodes, const int32_t* in_ptr, int32_t* ovf_ptr);

/* Fills f->groups and f->wsum for the batch (one launch).  eig is [n_nodes, n_eig] with row stride ld_eig. */
int dgn_field_build(const DgnGraph* g, const DgnAggSpec* spec, const float* eig, int32_t ld_eig, const DgnField* f,
                    void* stream);

/* Fused layer epilogue, rb/nets/dgn_layer.py:122-130:
 *   z = y * snorm_n ; BatchNorm1d(z) (batch statistics, running stats updated) ; ReLU ; + residual.
 * Two stream-ordered launches.  `stats` is a [DGN_NORM_WS_FLOATS(C)] fp32 workspace: on return it
 * holds mean[C], rstd[C] (needed by the backward) followed by scratch.  n_rows_dev (optional, device
 * int32) overrides n_rows so that padded batches can be replayed from a CUDA graph. */
typedef struct {
  int32_t n_rows, n_cols;
  const float* y;          int32_t ld_y;     /* posttrans output                                        */
  const float* y_bias;     /* [C] optional: z = (y + y_bias) * snorm_n (the posttrans bias, fused)      */
  const float* snorm;      /* [n_rows] graph-norm factor per node, NULL = graph_norm off                */
  const float* gamma;      /* [C] BatchNorm weight, NULL = batch_norm off                               */
  const float* beta;       /* [C] BatchNorm bias                                                        */
  float* running_mean;     /* [C] updated in place when training (NULL = skip)                          */
  float* running_var;      /* [C]                                                                       */
  float momentum, eps;
  int32_t training;        /* 1: batch statistics; 0: running statistics                                */
  int32_t relu;            /* 1: ReLU after the norm (complex/simple), 0: none (tower)                  */
  const float* residual;   int32_t ld_res;   /* NULL = no residual                                      */
  float* out;              int32_t ld_o;
  float* stats;            /* [DGN_NORM_WS_FLOATS(C)] workspace, see above                              */
  const int32_t* n_rows_dev;
  int32_t stat_parts;      /* > 0: `stats` already holds that many partial batch statistics written by
                              dgn_post_forward (DgnPostStats): the statistics pass is skipped (one launch)      */
} DgnNormArgs;

int dgn_norm_forward(const DgnNormArgs* a, void* stream);

/* dgn_norm_forward (apply pass only: a->stat_parts > 0, or no batch statistics needed) fused with the node-level halves
 * of the NEXT layer's pretrans: out = epilogue(y) as above, P = out W_src^T, Q = out W_dst^T with W = [W_src | W_dst | ..]
 * of shape [f_out, >= 2 C] - one launch instead of dgn_norm_forward + dgn_pair_linear_forward.  C, f_out <= 128 and
 * multiples of 4, else DGN_ERR_UNSUPPORTED. */
int dgn_norm_pair_forward(const DgnNormArgs* a, int32_t f_out, const float* w, int32_t ld_w, float* p, int32_t ld_p,
                          float* q, int32_t ld_q, void* stream);

/* Backward of dgn_norm_forward: given g_out it writes d_y, d_residual (may alias nothing; NULL to
 * skip), d_gamma, d_beta.  Needs the forward's `out` (for the ReLU mask) and `stats`. */
typedef struct {
  const float* g_out;      int32_t ld_go;
  float* d_y;              int32_t ld_dy;
  float* d_residual;       int32_t ld_dres;
  float* d_gamma;          /* [C] */
  float* d_beta;           /* [C] */
  float* d_bias;           /* [C] optional: gradient of y_bias = column sums of d_y                     */
  int32_t accumulate;      /* 1: d_gamma / d_beta / d_bias are accumulated (+=) instead of overwritten  */
  float* scratch;          /* [DGN_NORM_WS_FLOATS(C)] fp32 workspace                                    */
} DgnNormGrad;

int dgn_norm_backward(const DgnNormArgs* a, const DgnNormGrad* g, void* stream);

/* Per-graph readout over contiguous node segments: op 0 = sum, 1 = mean, 2 = max.
 * graph_ptr [n_graphs+1] device.  Replaces dgl.sum_nodes / mean_nodes / max_nodes. */
/* d_weight[idx[r], :] += g[r, :] for r < n_rows: gradient of an embedding lookup
 * (nn.Embedding at rb/nets/molecules_graph_regression/dgn_net.py:36,58).  Deterministic (fixed summation
 * order); vocab * 1 KiB of shared memory must fit (vocab <= 200).  idx is int64 (torch.long).
 * ws: DGN_EMB_WS_FLOATS(vocab, n_cols) floats of device workspace that must be ZERO before the first call
 * (the kernel leaves its counters zeroed again) and may be reused by later calls on the same stream. */
#define DGN_EMB_WS_FLOATS(V, C) (32 * (V) * (C) + 64)
int dgn_embedding_backward(int32_t n_rows, int32_t n_cols, int32_t vocab, const int64_t* idx, const float* g,
                           int32_t ld_g, float* d_weight, int32_t ld_w, const int32_t* n_rows_dev, float* ws,
                           void* stream);

/* One Adam update of a flat fp32 parameter buffer (torch.optim.Adam as used at rb/main_molecules.py:82:
 * weight decay added to the gradient, bias-corrected moments).  state: 2 device int32, zero-initialised:
 * state[0] = steps taken so far (incremented by the kernel), state[1] = internal.  All pointers 16 B aligned.
 * hyper (optional, DEVICE, 3 floats {lr, weight_decay, grad_scale}): when given it overrides the by-value lr /
 * weight_decay and multiplies the gradient by grad_scale (1 / world size after a sum all-reduce), so a launch that
 * was captured into a CUDA graph follows a learning-rate scheduler (ReduceLROnPlateau, rb/main_molecules.py:89-130). */
int dgn_adam_step(int64_t n, float* param, const float* grad, float* exp_avg, float* exp_avg_sq, float lr, float beta1,
                  float beta2, float eps, float weight_decay, const float* hyper, int32_t* state, void* stream);

/* Gradient all-reduce over NVLink peer memory fused with the Adam update (data-parallel training): one
 * launch per step and rank instead of an NCCL all-reduce + dgn_adam_step, capturable in the step's CUDA graph.  Every
 * rank's flat gradient buffer lives in symmetric (peer-mapped) device memory; the kernel reduce-scatters it in rank order
 * (deterministic, every rank gets bit-identical sums), then all-gathers the slices and applies Adam to the local replica
 * (grad_scale = 1 / world through `hyper`).  All ranks must launch it the same number of times.
 *   grad_ptrs  DEVICE array [world] of uint64: the ranks' gradient buffers as mapped into THIS process
 *   flag_ptrs  DEVICE array [world] of uint64: the ranks' signal pads, DGN_AR_FLAG_WORDS(world) uint32 each, zero before
 *              the first launch
 *   epoch      local DEVICE uint32[4], zero before the first launch; [2] becomes 1 if a peer did not reach a barrier
 *              within ~30 s (the step's result is then invalid; the kernel does not hang)
 * n must be a multiple of 4 floats (the engine pads its flat buffers); other arguments as dgn_adam_step. */
#define DGN_AR_BLOCKS 148
#define DGN_AR_MAX_WORLD 8
#define DGN_AR_FLAG_WORDS(world) (3 * DGN_AR_BLOCKS * (world))
typedef struct {
  int32_t world, rank;
  const uint64_t* grad_ptrs;
  const uint64_t* flag_ptrs;
  uint32_t* epoch;
  float* reduced;             /* optional local DEVICE buffer of n floats: receives the all-reduced SUM (checks, logging) */
  int32_t one_shot_max_world; /* world <= this: every rank reads all buffers in full (one barrier fewer); else
                                 reduce-scatter + all-gather.  2 is a good default for ~2 MB of gradients */
} DgnPeerGroup;
int dgn_allreduce_adam(const DgnPeerGroup* pg, int64_t n, float* param, float* exp_avg, float* exp_avg_sq, float lr,
                       float beta1, float beta2, float eps, float weight_decay, const float* hyper, int32_t* state,
                       void* stream);

/* C (+)= op(A) * op(B) in fp32 accuracy on the tcgen05 tensor cores (3xTF32 split, fp32 accumulation in tensor
 * memory).  Replaces the library GEMMs of the pre/post-transform MLPs (FCLayer, rb/nets/layers.py:76-100).
 *   a_kmajor = 1: A is [M][K] with row stride lda;  0: A is [K][M]
 *   b_kmajor = 1: B is [N][K] with row stride ldb;  0: B is [K][N]
 *   c_transposed = 1: the result is written as C[N][M] (row stride ldc);  accumulate = 1: C += result
 * All operand rows must be 16 B aligned with contiguous extents that are multiples of 4 floats, otherwise
 * DGN_ERR_UNSUPPORTED is returned (callers then use the library GEMM).  ws: dgn_gemm_ws_floats() floats of
 * device memory, zero before the first call; deterministic (split-K partials are added in a fixed order). */
int64_t dgn_gemm_ws_floats(void);
int dgn_gemm_tf32x3(int32_t M, int32_t N, int32_t K, const float* A, int32_t lda, int32_t a_kmajor, const float* B,
                    int32_t ldb, int32_t b_kmajor, float* C, int32_t ldc, int32_t accumulate, int32_t c_transposed,
                    float* ws, void* stream);

/* Scaler-folded posttrans of one DGN layer.  The reference concatenates the aggregates once per scaler,
 * cat = [h | c_0 agg | ... | c_{S-1} agg] (rb/nets/dgn_layer.py:94-96, :116) and applies posttrans to it (:119).  c_s is a
 * per-node factor (rb/nets/scalers.py:7-21), so with W = [W_h | W_0 | ... | W_{S-1}]
 *     y[v] = h[v] W_h^T + sum_s c_s(v) (agg[v] W_s^T)
 * and the [N, S*A*F] tensor never has to exist: `cat` here is [h | agg] with the RAW aggregates (dgn_agg_forward with a
 * single scaler), one fp32 accumulator per scaler lives in tensor memory and the epilogue folds them with c_s(v).
 *   dgn_post_forward   y = the expression above (no bias; dgn_norm_forward adds it as y_bias)
 *   dgn_post_backward  d_cat[:, :F] = d_y W_h ; d_cat[:, F + c] = sum_s c_s(v) (d_y W_s)[c]   (gradient of `cat`)
 *   dgn_post_wgrad     d_w (+)= [ d_y^T h | (c_0 d_y)^T agg | ... ]                             (gradient of `w`)
 * tcgen05 3xTF32 like dgn_gemm_tf32x3; split-K partial tiles are reduced across a thread-block cluster through
 * distributed shared memory (deterministic, no workspace).  All widths / leading dimensions multiples of 4 floats,
 * pointers 16 B aligned, else DGN_ERR_UNSUPPORTED. */
typedef struct {
  int32_t n_rows;          /* N (row capacity of a padded batch)                                               */
  int32_t n_lead;          /* F: leading columns of cat that no scaler touches (the copy of h; 0 = none)        */
  int32_t n_agg;           /* A*F: columns of raw aggregates                                                    */
  int32_t n_out;           /* F_out                                                                             */
  int32_t n_scalers;       /* S as declared; like rb/nets/dgn_layer.py:95 the scalers are applied only if S > 1 */
  uint8_t scaler_kind[DGN_MAX_SCALERS];
  float avg_log;           /* avg_d["log"]                                                                      */
  const float* log_deg;    /* [N] log(in_degree + 1) (DgnGraph.log_deg)                                         */
  const float* cat;        /* [N, n_lead + n_agg]                                                               */
  int32_t ld_cat;
  const float* w;          /* [n_out, n_lead + S*n_agg]  posttrans weight (nn.Linear layout)                    */
  int32_t ld_w;
} DgnPostArgs;

/* Optional by-product of dgn_post_forward: partial batch statistics (count, mean, M2 per column and per row slab) of
 * z = (y + y_bias) * snorm, i.e. of what BatchNorm1d normalises at rb/nets/dgn_layer.py:122-126, written into the
 * workspace of the dgn_norm_forward call that follows, one slab per 128-row tile (the CTAs of a cluster combine theirs
 * over distributed shared memory).  *stat_parts receives the number of slabs (pass it on as DgnNormArgs.stat_parts),
 * 0 when the layout does not fit the workspace (then run the norm's own statistics pass). */
typedef struct {
  float* stats;              /* DgnNormArgs.stats of the following dgn_norm_forward, [DGN_NORM_WS_FLOATS(n_out)] */
  const float* y_bias;       /* [n_out] or NULL                                                                */
  const float* snorm;        /* [N] or NULL                                                                    */
  const int32_t* n_rows_dev; /* device int32: real row count of a padded batch, or NULL                        */
} DgnPostStats;

int dgn_post_forward(const DgnPostArgs* a, float* y, int32_t ld_y, const DgnPostStats* st, int32_t* stat_parts,
                     void* stream);
int dgn_post_backward(const DgnPostArgs* a, const float* d_y, int32_t ld_dy, float* d_cat, int32_t ld_dcat, void* stream);
int dgn_post_wgrad(const DgnPostArgs* a, const float* d_y, int32_t ld_dy, float* d_w, int32_t ld_dw, int32_t accumulate,
                   void* stream);

/* Weight / bias gradient of the 1-layer pretrans split per node (rb/nets/dgn_layer.py:75-80), one launch:
 *   d_w[:, :Fi] (+)= d_P^T h ;  d_w[:, Fi:2Fi] (+)= d_Q^T h ;  d_b (+)= column sums of d_Q   (d_b may be NULL)
 * h [N, Fi], d_P / d_Q [N, Fo] with one common leading dimension, d_w [Fo, >= 2 Fi]. */
int dgn_pre_wgrad(int32_t n_rows, int32_t f_in, int32_t f_out, const float* h, int32_t ld_h, const float* d_p,
                  int32_t ld_dp, const float* d_q, int32_t ld_dq, float* d_w, int32_t ld_dw, float* d_b,
                  int32_t accumulate, void* stream);

/* Node-level halves of the 1-layer pretrans (rb/nets/dgn_layer.py:75-80) in one launch each, straight from the
 * parameter W = [W_src | W_dst | ...] of shape [Fo, ld_w >= 2*Fi]:
 *   forward : P = h W_src^T, Q = h W_dst^T            (h [N,Fi], P and Q [N,Fo])
 *   backward: d_h += d_P W_src + d_Q W_dst            (accumulates into d_h [N,Fi])
 * fp32 FMA on CUDA cores (these K <= 128 products are launch-latency bound).  Fi, Fo <= 128 and multiples of 4,
 * outputs 16 B aligned, else DGN_ERR_UNSUPPORTED. */
int dgn_pair_linear_forward(int32_t N, int32_t Fi, int32_t Fo, const float* h, int32_t ld_h, const float* W, int32_t ld_w,
                            float* P, int32_t ld_p, float* Q, int32_t ld_q, void* stream);
int dgn_pair_linear_backward(int32_t N, int32_t Fi, int32_t Fo, const float* dP, int32_t ld_p, const float* dQ,
                             int32_t ld_q, const float* W, int32_t ld_w, float* d_h, int32_t ld_dh, void* stream);

/* dgn_pair_linear_backward with the source-side reduction of dgn_agg_backward folded in: call dgn_agg_backward with
 * DgnAggGrad.d_x = NULL and edge_ws set (the per-edge message gradients are left in edge_ws, slot order, and the
 * source-side launch is skipped), then
 *   d_P[u] = sum over out-edges j of u of edge_ws[out_slot[j]] ;  d_h[u] += d_P[u] W_src + d_Q[u] W_dst
 * in ONE launch; d_P is written out (the weight gradient needs it).  out_ptr / out_slot: DgnGraph's by-source transpose. */
int dgn_pair_gather_backward(int32_t N, int32_t Fi, int32_t Fo, const int32_t* out_ptr, const int32_t* out_slot,
                             const float* edge_ws, int32_t ld_ws, const float* dQ, int32_t ld_q, const float* W,
                             int32_t ld_w, float* d_h, int32_t ld_dh, float* dP, int32_t ld_p, void* stream);

/* MLPReadout with L = 2 hidden layers (rb/nets/mlp_readout_layer.py:11-30):
 *   y = W3 relu(W2 relu(W1 x + b1) + b2) + b3      x [n_rows, d0], W1 [d1, d0], W2 [d2, d1], W3 [d_out, d2]
 * in ONE launch (one CTA, weights resident in shared memory).  Weights are contiguous row-major ([out][in], the
 * nn.Linear layout).  a1 [n_rows, d1] and a2 [n_rows, d2] receive the post-ReLU activations the backward needs.
 * Limits (else DGN_ERR_UNSUPPORTED, callers use the library): n_rows <= DGN_HEAD_MAX_ROWS, d1*d0 <= 4096,
 * d2*d1 <= 1024, d_out*d2 <= 1024, d1 + d2 + d_out <= 1024. */
#define DGN_HEAD_MAX_ROWS 1024
typedef struct {
  int32_t n_rows, d0, d1, d2, d_out;
  const float* x;  int32_t ld_x;
  const float* w1; const float* b1;
  const float* w2; const float* b2;
  const float* w3; const float* b3;
  float* a1;       /* [n_rows, d1] contiguous */
  float* a2;       /* [n_rows, d2] contiguous */
  float* y;        int32_t ld_y;
} DgnHeadArgs;

/* Gradients of dgn_head_forward given g_y [n_rows, d_out]; any output may be NULL.  accumulate = 1 adds into the
 * weight / bias gradient buffers (d_x is always overwritten).  Deterministic: fixed owner thread per entry, rows in order. */
typedef struct {
  const float* g_y; int32_t ld_gy;
  float* d_x;       int32_t ld_dx;
  float* d_w1; float* d_b1;
  float* d_w2; float* d_b2;
  float* d_w3; float* d_b3;
  int32_t accumulate;
} DgnHeadGrad;

int dgn_head_forward(const DgnHeadArgs* a, void* stream);
int dgn_head_backward(const DgnHeadArgs* a, const DgnHeadGrad* g, void* stream);

/* loss[0] = mean |y[i] - target[i]| over n contiguous elements (nn.L1Loss, reduction = mean), one launch, fixed
 * summation order; backward: d_y[i] = g_loss[0] * sign(y[i] - target[i]) / n with sign(0) = 0. */
int dgn_l1_loss_forward(int32_t n, const float* y, const float* target, float* loss, void* stream);
int dgn_l1_loss_backward(int32_t n, const float* y, const float* target, const float* g_loss, float* d_y, void* stream);

/* Device-side collation: dgl.batch + dataset.collate (rb/data/molecules.py:219-230) without the host.
 * The whole dataset lives in HBM as ONE pre-batched graph ("fragments"): node / edge ranges per graph, the CSR of the
 * giant block-diagonal batch, per-node log-degree and overflow-group counts, and the payload arrays (node features,
 * eigenvectors, edge features, per-graph targets).  A mini-batch is then an index list: one launch copies the selected
 * graphs' fragments behind each other into the fixed-capacity batch layout the kernels read (BatchedGraph), adding the
 * node / edge offsets on the fly; snorm_n, graph_ptr, ovf_ptr, meta and the zero padding are produced in the same
 * launch.  Bit-identical to dgn_build_csr_host + dgn_build_groups_host on the concatenated edge lists (block-diagonal
 * batches sort per graph). */
#define DGN_MAX_PAYLOADS 6
typedef struct {
  const void* src;        /* dataset array, rows of row_bytes (multiple of 4)                                     */
  void* dst;              /* batch array of the same row size; padding rows are zeroed                           */
  int32_t row_bytes;
  int32_t per;            /* 0: per node, 1: per edge, 2: per graph                                              */
} DgnPayload;
typedef struct {
  int32_t n_graphs;       /* graphs in the dataset                                                               */
  const int32_t* node_off;/* [G+1] node range of every graph in the dataset arrays                               */
  const int32_t* edge_off;/* [G+1] edge range                                                                    */
  const int32_t* ovf_off; /* [G+1] overflow-group range (dgn_build_groups_host over the whole dataset)           */
  const int32_t* in_ptr;  /* [Nt+1] CSR of the whole dataset as one batch (dgn_build_csr_host), global ids       */
  const int32_t* in_src;  /* [Et]                                                                                */
  const int32_t* in_eid;  /* [Et]                                                                                */
  const int32_t* out_ptr; /* [Nt+1]                                                                              */
  const int32_t* out_slot;/* [Et]                                                                                */
  const int32_t* src;     /* [Et] edge list, global ids                                                          */
  const int32_t* dst;     /* [Et]                                                                                */
  const float* log_deg;   /* [Nt]                                                                                */
  const int32_t* ovf_ptr; /* [Nt+1] dataset-wide overflow-group offsets                                          */
} DgnDataset;
typedef struct {
  int32_t n_cap, e_cap, b_cap;   /* capacities of the batch layout (nodes, edges, graphs)                        */
  int32_t *in_ptr, *in_src, *in_eid, *out_ptr, *out_slot, *src, *dst, *graph_ptr, *ovf_ptr, *meta;
  float *log_deg, *snorm_n;
  int32_t n_payloads;
  DgnPayload payload[DGN_MAX_PAYLOADS];
} DgnBatchOut;
/* ids: DEVICE int32 [n_ids] graph indices in batch order.  Returns DGN_ERR_INVALID for bad arguments; a batch that
 * exceeds the capacities is truncated on the device and flagged in meta[3] = 1 (callers check it lazily). */
int dgn_collate_device(const DgnDataset* ds, const int32_t* ids, int32_t n_ids, const DgnBatchOut* out, void* stream);

/* One launch that copies n_segments rectangular fp32 blocks between two sets of tensors, described by a DEVICE table of
 *   struct { float* a; float* b; int32_t rows, cols, ld_a, ld_b; }           (32 bytes per entry)
 * direction 0: b = a; 1: a = b; 2: a += b.  Used to pack the per-tower parameters of DGNLayerTower
 * (rb/nets/dgn_layer.py:279-307) into the dense block-structured operands of the single-launch tower path and to scatter
 * the gradients / running statistics back. */
int dgn_segment_copy(const void* seg_table, int32_t n_segments, int32_t direction, void* stream);

/* Laplacian eigenvectors of every graph on the device: the per-graph host loop of the reference's loaders
 * (scipy.sparse.linalg.eigs on L, rb/data/molecules.py:100-116, rb/data/SBMs.py:110-139, rb/data/HIV.py:17-46).
 * node_off [G+1] node range of every graph; in_ptr / in_src: CSR with GLOBAL node ids (DgnDataset / DgnGraph arrays);
 * the adjacency is symmetrised, deg = clip(row sum, 1).  norm: 0 = 'none' (L = D - A), 1 = 'sym' (I - D^-1/2 A D^-1/2),
 * 2 = 'walk' (I - D^-1 A).  eig [N_total, ld_eig] receives the k eigenvectors of smallest eigenvalue (ascending, unit
 * norm, largest-magnitude entry positive; zero columns for graphs with fewer than k nodes), eigval [G, k] (optional)
 * the eigenvalues.  One CTA per graph, one-sided Jacobi in shared memory: max_nodes (the largest graph) <= 238. */
int dgn_eig_precompute(int32_t n_graphs, const int32_t* node_off, const int32_t* in_ptr, const int32_t* in_src,
                       int32_t max_nodes, int32_t norm, int32_t k, float* eig, int32_t ld_eig, float* eigval, void* stream);
/* The sign-flip augmentation of rb/train/train_molecules_graph_regression.py:29-33 on the device: every ENTRY of eig
 * is negated with probability 1/2 (counter-based generator over (seed, step, index): replayable from a CUDA graph when
 * `step` lives in the caller's loop). */
int dgn_eig_flip(float* eig, int64_t n_elems, uint64_t seed, uint64_t step, void* stream);

int dgn_readout_forward(int32_t n_graphs, const int32_t* graph_ptr, int32_t n_cols, const float* h, int32_t ld_h,
                        int32_t op, float* out, int32_t ld_o, void* stream);
/* d_h has n_rows_total rows: rows past graph_ptr[n_graphs] (padding of a fixed-capacity batch) are zeroed. */
int dgn_readout_backward(int32_t n_graphs, const int32_t* graph_ptr, int32_t n_cols, const float* h, int32_t ld_h,
                         const float* out, int32_t ld_o, int32_t op, const float* g_out, int32_t ld_go,
                         float* d_h, int32_t ld_dh, int32_t n_rows_total, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* DGN_B200_H_ */
