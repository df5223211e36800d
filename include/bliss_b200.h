/* bliss_b200.h — C ABI of the B200-native BLISS sample-and-aggregate hot path.
 *
 * Every entry point replaces a group of DGL FFI calls / torch glue the reference reaches from
 * Python (the reference has no native code of its own; citations are into /root/reference):
 * raw device pointers, element counts, scalars and a CUDA stream go in; an int comes back
 * (0 = ok, <0 = bad argument, >0 = cudaError_t of the launch).  No entry point allocates,
 * synchronises or keeps global state: the caller owns every buffer (sizes below), data-dependent
 * sizes are written to the device-side bliss_counters block and read back by the caller when it
 * needs them.  All kernels run on `stream` (pass torch's current stream).
 *
 * Data layout: the graph is CSC — indptr int64 [V+1], indices int32 [E] (source
 * of each in-edge, column = destination, self-loop last in its column), per-edge arrays
 * (EXP3 weights, static weights `w`) are stored in CSC order so they stream with `indices`.
 */
#ifndef BLISS_B200_H
#define BLISS_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define BLISS_B200_VERSION 100

/* importance modes of bliss_frontier_prob */
#define BLISS_MODE_BANDIT 0   /* q_ij from EXP3 weights, p_j = sqrt(sum_i (q_ij/sum_k q_ik)^2)  bandit_sampler.py:47-82,101-138 */
#define BLISS_MODE_LADIES 1   /* p_j = sqrt(sum_i w_ij^2) with static w                         ladies_sampler.py:34-52 */
#define BLISS_MODE_UNIFORM 2  /* flag OR-ed onto the above: importance_sampling=0, p_j = 1 for nodes with an out-edge  bandit_sampler.py:77-81 */

#define BLISS_MODE_NEIGHBOR 4 /* flag OR-ed onto BLISS_MODE_LADIES: uniform neighbour sampling — every seed keeps min(fanout,
                                degree) in-edges chosen uniformly without replacement (fanout <= 0: all): DGL NeighborSampler /
                                MultiLayerFullNeighborSampler, train_lightning.py:349-357.  No probabilities, no node selection. */
#define BLISS_MODE_PLANNED 8  /* flag for bliss_sample_layer_front: bliss_frontier_plan for this layer (same seeds, same
                                workspace) was already enqueued — the plan depends on the seeds only, so a caller that knows
                                them early (the top layer's: the batch) can run it before the weights it samples with are final */

#define BLISS_COLLECT_BITMAP 16 /* flag OR-ed onto the mode: collect candidates from a bitmap the scatter marks
                                  (sparse frontiers in huge graphs) instead of a dense scan of the |V| accumulators */

/* aggregation modes of bliss_spmm */
#define BLISS_AGG_SUM 0
#define BLISS_AGG_MEAN 1      /* divide by max(in_degree, 1)  (fn.mean, dglnn.SAGEConv 'mean') */

typedef struct bliss_graph {
  int64_t num_nodes;
  int64_t num_edges;
  const int64_t* indptr;   /* [num_nodes + 1] */
  const int32_t* indices;  /* [num_edges]     */
  const int32_t* eid;      /* [num_edges] CSC position -> original edge id (may be NULL) */
} bliss_graph;

/* Device-resident per-layer counters; the host reads this block back once per layer. */
typedef struct bliss_counters {
  int32_t n_seeds;     /* rows of this layer                                     */
  int32_t n_cand;      /* |seeds ∪ sources|  (insg.num_nodes(), :392)            */
  int32_t n_sel;       /* selected non-seed candidates                           */
  int32_t n_src;       /* block source nodes = n_seeds + n_sel                   */
  int32_t n_heavy;     /* reserved (0)                                           */
  int32_t n_light;     /* reserved (0)                                           */
  int32_t take_all;    /* 1 when n_cand <= fanout (:392-393)                     */
  int32_t iters;       /* scale-search iterations used (:396)                    */
  int64_t e_in;        /* in-edges of the seeds (insg.num_edges())               */
  int64_t n_edges;     /* block edges E_b                                        */
  double  c;           /* Poisson scale                                          */
  double  s_last;      /* last S = sum min(c p, 1)                               */
  int32_t queue[8];    /* work-queue cursors of the chunk passes (prob 1-3, count, fill) */
  int32_t error;       /* non-zero: a capacity was exceeded (see BLISS_ERR_*)    */
  int32_t n_chunks;    /* 256-edge warp-chunks of the frontier's rows            */
} bliss_counters;

#define BLISS_ERR_SEL_CAPACITY 1
#define BLISS_ERR_EDGE_CAPACITY 2

/* Per-sampler workspace (device pointers; V = num_nodes, S = max seeds of a layer,
 * C = capacity of selected nodes).  Invariant between layers: acc == 0, first_pos == ~0,
 * node_info[v].x == -1, sel_bits == cand_bits == 0 — bliss_block_finish restores it for every node a layer
 * touched, so nothing |V|-sized is cleared per step. */
typedef struct bliss_workspace {
  uint64_t* acc;        /* [V]  fixed-point column accumulator (sum of squared edge terms)  */
  uint64_t* first_pos;  /* [V]  first occurrence key of a selected source                 */
  int32_t*  node_info;  /* [2V] (local id | -1, float bits of inclusion prob P) per node  */
  uint32_t* sel_bits;   /* [(V+31)/32] bitmap: node is selected                           */
  uint32_t* cand_bits;  /* [(V+31)/32] candidate bitmap (BLISS_COLLECT_BITMAP mode only)    */
  uint32_t* keep_bits;  /* [8 * (E/256+V)] one bit per edge of every chunk: edge kept in the block */
  int32_t*  cand;       /* [V]  candidate list: seeds first, then sources unordered       */
  float*    p_cand;     /* [V]  raw probability per candidate slot                        */
  int32_t*  sel;        /* [C]  selected non-seed candidates, unordered                   */
  int64_t*  row_a;      /* [S]  CSC start of every seed's column (by seed rank)           */
  int32_t*  row_d;      /* [S]  in-degree of every seed                                   */
  int32_t*  chunk_first;/* [S+1] first 256-edge chunk of every seed's column (prefix)     */
  void*     chunk_rec;  /* [E/256+V] 32-byte record per chunk (CSC start, length, row, degree, offset, chunk range) */
  double*   part_w;     /* [E/256+V] per-chunk partial of sum_j w_ij                      */
  double*   part_q;     /* [E/256+V] per-chunk partial of sum_j q_ij                      */
  float*    row_w;      /* [S]  sum_j w_ij  per seed                                      */
  float*    row_q;      /* [S]  sum_j q_ij  per seed                                      */
  int32_t*  row_cnt;    /* [S]  kept in-edges per seed                                    */
  int32_t*  part_cnt;   /* [E/256+V] kept in-edges per chunk                              */
  double*   part_t;     /* [E/256+V] per-chunk partial of the unnormalised block weights  */
  int32_t*  chunk_pre;  /* [E/256+V] kept in-edges of the chunk's row before the chunk    */
  float*    row_t;      /* [S]  sum of the unnormalised block weights per seed            */
  int64_t   cap_seeds;  /* S */
  int64_t   cap_sel;    /* C */
  bliss_counters* ctr;  /* one counters block                                             */
  /* sync-free operation (CUDA-graph replay): values the kernels read from the device instead of
   * taking them from the host arguments, which then only give capacities.  NULL = use the host value. */
  const int32_t*  n_seeds_dev;  /* true seed count of this layer (e.g. the previous layer's n_src) */
  const uint64_t* step_dev;     /* Philox step counter                                            */
  /* Optional host-visible copy of the layer's counters (pinned, device-mapped host memory): written by
   * bliss_block_finish, the layer's last kernel, so a replayed step needs no separate device->host copy. */
  bliss_counters* ctr_mirror;
} bliss_workspace;

/* Outputs of one sampled layer (device pointers, capacities checked against the counters). */
typedef struct bliss_block_out {
  int32_t* indptr;      /* [n_seeds + 1] destination-major CSR                            */
  int32_t* edge_src;    /* [E_b] local source id                                          */
  int32_t* edge_dst;    /* [E_b] local destination id                                     */
  int64_t* csc_pos;     /* [E_b] position of the edge in the graph CSC                    */
  int32_t* eid;         /* [E_b] original edge id (NULL to skip)                          */
  float*   q_ij;        /* [E_b] edge probability q_ij (NULL for LADIES)                  */
  float*   edge_w;      /* [E_b] block weight W~ (bandit_sampler.py:314-320)              */
  int32_t* src_nid;     /* [n_src] global id of each block source                         */
  float*   node_prob;   /* [n_src] inclusion probability P (bandit_sampler.py:328)        */
  int32_t* out_deg;     /* [n_src] block out-degree of each source (NULL to skip)         */
  uint32_t* t_bits;     /* [cap_src x t_words] source x destination bitmap of the block, marked by the
                           fill for bliss_block_transpose (NULL to skip)                  */
  int64_t  t_words;     /* row stride of t_bits in 32-bit words (>= ceil(n_seeds / 32))   */
  int32_t* seg_ptr;     /* [n_seeds+1] prefix of max(1, ceil(in-degree/32)): the 32-edge row segments the
                           balanced SpMM works on (padded rows count one segment each; NULL ok)       */
  float*   inv_deg;     /* [n_seeds] 1 / max(block in-degree, 1)  (fn.mean divisor; NULL ok)  */
  int64_t  cap_edges;
  int64_t  cap_src;
  int64_t  pad_src;     /* > n_src: out_deg[n_src..pad_src) = 0 (capacity padding)                */
  int64_t  pad_rows;    /* > n_seeds: indptr[n_seeds+1..pad_rows] = E_b, inv_deg[n_seeds..pad_rows) = 1
                           (capacity-padded blocks for CUDA-graph replay); 0 = no padding          */
} bliss_block_out;

/* Peer-memory exchange of the sparse bandit updates (data parallel, one node): every rank owns a window in symmetric
 * memory, mapped by all ranks:  [2 parities][world slots, one per source rank] ++ flags[2][n_layers][world] (uint64);
 * a slot is laid out like the packed exchange buffer (int64 edge counts, then per layer int32 pos[cap] | fp32 x[cap]).
 * bliss_reward_update with a bliss_p2p writes the layer's (position, exponent) pairs into ITS slot of EVERY window
 * (NVLink stores) and raises flag = *step_dev + 1 in every window when its last CTA is done; bliss_apply_updates_p2p
 * waits for all ranks' flags of the layer and applies the slots of the rank's own window.  parity = *step_dev & 1. */
typedef struct bliss_p2p {
  const int64_t* peer_base;   /* device array [world]: address of every rank's window in this rank's address space */
  int32_t  world, rank;
  int64_t  parity_stride;     /* bytes between the two parity halves of a window (= world * rank_stride) */
  int64_t  rank_stride;       /* bytes of one slot */
  int64_t  count_off, pos_off, x_off;   /* this layer's offsets inside a slot (bytes) */
  int64_t  flags_off;         /* byte offset of the flags inside a window */
  int32_t  layer, n_layers;
  const int64_t* step_dev;    /* exchange step counter on the device */
  uint32_t* done_ctr;         /* [1] zero between launches: last-CTA detection of the producing kernel */
  int32_t  pull;              /* 0: the producer stores its slot into EVERY rank's window (push);
                                 1: it stores only into its OWN window and raises the flags everywhere, the consumer
                                    reads slot r from rank r's window over NVLink (pull: no duplicated stores) */
  int32_t  pad_;
  uint64_t mc_base;           /* 0, or the NVSwitch multicast address of the windows (same layout): one multimem.st
                                 per value reaches every rank's window, the switch replicates it (push mode only) */
} bliss_p2p;

/* Gradient all-reduce through peer memory, fused into the optimizer step: window = [2 parities][world slots of
 * slot_bytes] ++ flags[2][world] (uint64).  bliss_grad_push stores the flat gradient into the rank's slot of every
 * window and raises flag = *step_dev + 1; bliss_adam_step_p2p waits for all flags, adds the world slots of the own
 * window in rank order (bit-identical on every rank), divides by world and applies Adam (+ clears the gradients). */
typedef struct bliss_grad_p2p {
  const int64_t* peer_base;   /* device array [world] of window addresses */
  int32_t  world, rank;
  int64_t  parity_stride;     /* = world * slot_bytes */
  int64_t  slot_bytes;        /* >= 4 * n, multiple of 16 */
  int64_t  flags_off;
  const int64_t* step_dev;    /* exchange step counter (parity, flag value) */
  uint32_t* done_ctr;         /* [1], zero between launches */
  uint64_t mc_base;           /* 0, or the multicast address of the windows (see bliss_p2p) */
} bliss_grad_p2p;

int bliss_version(void);

/* Fill the workspace invariant (once, after allocation). */
int bliss_workspace_init(const bliss_workspace* ws, int64_t num_nodes, void* stream);

/* ---- (1) layer-importance probabilities ------------------------------------------------
 * replaces dgl.in_subgraph + compact_graphs + 3x copy_e_sum + e_div_v/e_div_u/v_add_e
 * (bandit_sampler.py:123-137, :67-75; ladies_sampler.py:42-48).                            */
int bliss_frontier_plan(const bliss_graph* g, const int32_t* seeds, int32_t n_seeds,
                        const bliss_workspace* ws, void* stream);
int bliss_frontier_prob(const bliss_graph* g, const int32_t* seeds, int32_t n_seeds,
                        const float* edge_weight_csc, float eta, int32_t mode,
                        const bliss_workspace* ws, void* stream);

/* ---- (2) inclusion probabilities + selection ---------------------------------------------
 * Poisson: scale search bandit_sampler.py:391-406 on the device (no host round trips), then
 * u < P with u = Philox4x32-10(key = seed, ctr = (nid, layer, step)) or u_inject[nid].
 * Top-k: torch.multinomial(prob, k, replacement=False) == topk(p / -log1p(-u)), :84-99.    */
int bliss_poisson_scale(int32_t n_seeds, int32_t fanout, double eps, int32_t poisson,
                        const bliss_workspace* ws, void* stream);
int bliss_select_poisson(int32_t n_seeds, uint64_t seed, uint64_t step, uint32_t layer,
                         const float* u_inject, const bliss_workspace* ws, void* stream);
/* scale search + selection fused in one thread-block-cluster launch (the Poisson fast path) */
int bliss_poisson_select(int32_t n_seeds, int32_t fanout, double eps, uint64_t seed, uint64_t step,
                         uint32_t layer, const float* u_inject, const bliss_workspace* ws,
                         void* stream);
int bliss_select_topk(int32_t n_seeds, int32_t fanout, uint64_t seed, uint64_t step, uint32_t layer,
                      const float* u_inject, float* key_scratch, const bliss_workspace* ws,
                      void* stream);
int bliss_philox_fill(uint64_t seed, uint64_t step, uint32_t layer, const int32_t* nids,
                      int64_t n, float* out, void* stream);   /* test hook */
/* uniform neighbour sampling of a planned frontier (after bliss_frontier_plan): per in-edge key =
 * word 0 of Philox4x32-10(key = seed, ctr = (CSC position, layer | 0x8000, step)); a seed with more than `fanout`
 * in-edges keeps its `fanout` smallest keys (ties in CSC order), else all (fanout <= 0: all).  Writes the keep bits and
 * registers the kept edges' sources; follow with bliss_block_count(mode | BLISS_MODE_NEIGHBOR), index, fill. */
int bliss_neighbor_select(const bliss_graph* g, int32_t n_seeds, int32_t fanout, uint64_t seed, uint64_t step,
                          uint32_t layer, const bliss_workspace* ws, void* stream);

/* ---- (3) block construction ---------------------------------------------------------------
 * replaces insg.subgraph + edge_subgraph + to_block + e_div_u/copy_e_sum/e_mul_v
 * (bandit_sampler.py:285-337; ladies_sampler.py:81-106).                                    */
int bliss_block_count(const bliss_graph* g, const int32_t* seeds, int32_t n_seeds,
                      const float* edge_weight_csc, float eta, int32_t mode,
                      const bliss_workspace* ws, void* stream);
int bliss_block_index(const int32_t* seeds, int32_t n_seeds, const bliss_workspace* ws,
                      const bliss_block_out* out, void* stream);
int bliss_block_fill(const bliss_graph* g, const int32_t* seeds, int32_t n_seeds,
                     const float* edge_weight_csc, float eta, int32_t mode,
                     const bliss_workspace* ws, const bliss_block_out* out, void* stream);
int bliss_block_finish(int32_t n_seeds, int32_t mode, const bliss_workspace* ws,
                       const bliss_block_out* out, void* stream);
/* The two calls of the host fast path: the same kernels as the stage entry points above, launched
 * back to back (front: plan, probabilities, selection, count, index; back: fill, finish). */
int bliss_sample_layer_front(const bliss_graph* g, const int32_t* seeds, int32_t n_seeds,
                             const float* edge_weight_csc, float eta, int32_t mode, int32_t fanout,
                             double eps, int32_t poisson, uint64_t seed, uint64_t step, uint32_t layer,
                             const float* u_inject, float* key_scratch /* top-k only */,
                             const bliss_workspace* ws, const bliss_block_out* out, void* stream);
int bliss_sample_layer_back(const bliss_graph* g, const int32_t* seeds, int32_t n_seeds,
                            const float* edge_weight_csc, float eta, int32_t mode,
                            const bliss_workspace* ws, const bliss_block_out* out, void* stream);
/* source-major transpose of a block (backward SpMM): t_indptr[n_src+1], t_dst[E], t_perm[E]
 * (edge ids ascending inside every source row, so backward sums are deterministic).  Driven by the
 * source x destination bitmap t_bits [n_src x t_words] (+ t_pre of the same shape, word prefixes):
 * have_counts = 0: out-degrees and bits are computed here from the edge list;
 * have_counts = 1: t_cursor holds the out-degrees and t_bits the marks (bliss_block_out.out_deg / t_bits). */
int bliss_block_transpose(const int32_t* edge_src, const int32_t* edge_dst, int64_t n_edges,
                          int32_t n_src, int32_t n_dst, int32_t* t_indptr, int32_t* t_cursor /* [n_src] */,
                          uint32_t* t_bits, int32_t* t_pre, int64_t t_words, int32_t* t_dst, int32_t* t_perm,
                          int32_t* t_seg_ptr /* [n_src+1] 32-edge segment prefix of the source rows; may be NULL */,
                          int32_t have_counts,
                          const int64_t* n_edges_dev /* true edge count on the device, or NULL */,
                          const float* edge_w /* [E] block weights, or NULL */,
                          float* t_w /* [E] the same weights in transpose order (t_w[k] = edge_w[t_perm[k]]), or NULL */,
                          void* stream);

/* ---- (5) aggregation ------------------------------------------------------------------------
 * replaces DGL g-SpMM u_mul_e/sum (+ fn.mean), g-SDDMM u_add_v, edge_softmax, th.norm and the
 * lazy feature gather (model.py:82-98,318-329,425-436; train_lightning.py:138).              */
int bliss_gather_rows(const float* table, const int32_t* nid, int64_t n_rows, int32_t dim,
                      float* out, float* row_norm /* may be NULL */, void* stream);
int bliss_row_norm(const float* x, int64_t n_rows, int32_t dim, float* out, void* stream);
/* y[i,:] = dscale_i * sum_{e in row i} w[perm? perm[e] : e] * sscale[col[e]] * x[col[e], :]
 * seg_ptr == NULL: one warp per row (whole-graph inference over short rows).
 * seg_ptr != NULL: [n_rows+1] prefix of max(1, ceil(len/32)) — rows are cut into 32-edge segments, a
 * CTA takes 8 consecutive segments (one per warp) and adds the runs that belong to one row in shared
 * memory; rows that span several CTAs get one partial per run and a second launch adds them in order
 * (deterministic, no atomics).  partial = scratch of item_cap * ceil(dim/tile)*tile floats
 * (item_cap >= seg_ptr[n_rows]; tile = the column tile, <= 1024). */
int bliss_spmm(const int32_t* indptr, const int32_t* col, const int32_t* perm, const float* w,
               const float* sscale, const float* dscale, int32_t agg, const float* x,
               int32_t n_rows, int32_t dim, const int32_t* seg_ptr, float* partial, int64_t item_cap,
               float* y, void* stream);
int bliss_gatv2_fwd(const int32_t* indptr, const int32_t* col, const float* feat /* [n_src,H,D] */,
                    const float* attn /* [H,D] */, const float* drop_mask /* [E,H] or NULL */,
                    float negative_slope, int32_t n_dst, int32_t heads, int32_t dim,
                    float* out /* [n_dst,H,D] */, float* logits /* [E,H] */,
                    float* row_max /* [n_dst,H] */, float* row_sum /* [n_dst,H] */, void* stream);
int bliss_gatv2_bwd_dst(const int32_t* indptr, const int32_t* col, const float* feat, const float* attn,
                        const float* drop_mask, const float* logits, const float* row_max,
                        const float* row_sum, const float* out, const float* grad_out,
                        float negative_slope, int32_t n_dst, int32_t heads, int32_t dim,
                        float* grad_logit /* [E,H] */, float* grad_feat /* [n_src,H,D], dst part */,
                        float* grad_attn /* [H,D], accumulated */, void* stream);
int bliss_gatv2_bwd_src(const int32_t* t_indptr, const int32_t* t_dst, const int32_t* t_perm,
                        const float* feat, const float* attn, const float* drop_mask,
                        const float* logits, const float* row_max, const float* row_sum,
                        const float* grad_out, const float* grad_logit, float negative_slope,
                        int32_t n_src, int32_t n_dst, int32_t heads, int32_t dim,
                        float* grad_feat /* in: dst-term rows [0,n_dst) from bwd_dst; out: full */,
                        void* stream);

/* ---- (4) bandit reward / weight update -----------------------------------------------------
 * replaces calculate_alpha / calculate_rewards / update_exp3_weights
 * (bandit_sampler.py:140-249).  alpha_mode 0: alpha = static w (SAGE/GCN); 1: GAT (a_ij, q_ij). */
int bliss_gat_alpha_sums(const int32_t* blk_indptr, const float* a_ij, const float* q_ij,
                         int32_t n_dst, float* asum, float* qsum, void* stream);
int bliss_reward_update(const bliss_graph* g, const int32_t* blk_indptr, const int32_t* edge_src,
                        const int32_t* edge_dst, const int64_t* csc_pos, const int32_t* dst_nid,
                        const float* q_ij, const float* node_prob, const float* embed_norm,
                        const float* w_static_csc, const float* a_ij, const float* asum,
                        const float* qsum, int32_t alpha_mode, float delta, int32_t n_dst,
                        int64_t n_edges, float* exp3_w_csc /* NULL: only emit */,
                        float* rewards /* [E_b] or NULL */,
                        float* x_out /* [E_b] clamped exponent, or NULL */,
                        double* l1_delta /* [1] accumulated sum(w_new - w_old), or NULL */,
                        const int64_t* n_edges_dev /* true edge count on the device (n_edges = capacity), or NULL */,
                        int64_t* count_out /* where to store the edge count (exchange header), or NULL */,
                        int32_t* pos_out /* [E_b] CSC positions as int32 (exchange send buffer; |E| < 2^31), or NULL */,
                        const bliss_p2p* p2p /* also store (pos, x) into every rank's window + raise flags, or NULL */,
                        float* wmax /* [1] running max of the updated weights (atomic max; weights are positive) — the
                                       range guard of the lazy normalisation re-scales when it nears FLT_MAX —, or NULL */,
                        void* stream);
/* wait for every rank's flag of p2p->layer (one polling CTA; a peer that never arrives sets bit `layer` of *error
 * after ~4 s instead of hanging), then w[pos] *= exp(x) for all ranks' slots of this rank's window */
int bliss_apply_updates_p2p(const bliss_p2p* p2p, int64_t cap, float* exp3_w_csc, double* l1_delta,
                            int32_t* error /* device int, or NULL */, float* wmax /* see bliss_reward_update */,
                            void* stream);
/* apply gathered updates from other ranks: w[pos[k]] *= exp(x[k]) */
int bliss_apply_updates(const int64_t* pos, const float* x, int64_t n, float* exp3_w_csc,
                        double* l1_delta, float* wmax, void* stream);
/* data-parallel fast path: apply all ranks' updates of one layer straight from the all-gathered
 * exchange buffer (per rank: int64 count at count_off, int32 pos[cap] at pos_off, fp32 x[cap] at
 * x_off; offsets in bytes).  The reward kernel wrote pos and x straight into the send buffer — no
 * packing pass, no host-side sizes; 8 bytes per sampled edge on the wire. */
int bliss_apply_updates_packed(const void* recv, int64_t rank_stride_bytes, int32_t world,
                               int64_t count_off, int64_t pos_off, int64_t x_off, int64_t cap,
                               float* exp3_w_csc, double* l1_delta, float* wmax, void* stream);
/* literal F.normalize(p=1) of one layer's weights (bandit_sampler.py:249): two launches. */
int bliss_l1_norm(const float* w, int64_t n, double* partial /* [1024] */, double* out /* [1] */,
                  void* stream);
int bliss_scale_by_inv(float* w, int64_t n, const double* norm, double eps, void* stream);


/* ---- hidden-layer epilogue (SAGE) ------------------------------------------------------------
 * replaces `fc_self(h_dst) + h_neigh` bias add, `activation` = relu, `dropout` (model.py:321-332) and the
 * next layer's `th.norm(h, dim=1)` (model.py:318) — and their backward passes plus the bias-gradient
 * reduction — with one launch each way.  dim % 4 == 0, dim <= 1024, 16-byte aligned rows.
 * Dropout: u = Philox4x32-10(key = seed, ctr = (element / 4, layer, step)) words -> [0,1); kept iff
 * u < 1 - p, scaled by 1 / (1 - p); *step_dev is read on the device (CUDA-graph replay). */
int bliss_sage_epilogue_parts(void);   /* rows of the bias_partial scratch (= CTAs of the backward launch) */
int bliss_sage_epilogue_fwd(const float* a, const float* b, const float* bias /* [dim] or NULL */,
                            int32_t n_rows, int32_t dim, int32_t relu, float p_drop, uint64_t seed,
                            const int64_t* step_dev, uint32_t layer, float* y,
                            float* row_norm /* [n_rows] or NULL */, void* stream);
int bliss_sage_epilogue_bwd(const float* grad_y, const float* y, int32_t n_rows, int32_t dim,
                            int32_t gate /* 1: grad_z = grad_y * [y > 0] / (1 - p); 0: grad_z = grad_y */,
                            float p_drop, float* grad_z, float* bias_partial /* [parts, dim] or NULL */,
                            float* grad_bias /* [dim] or NULL */, void* stream);

/* ---- loss ------------------------------------------------------------------------------------
 * replaces nn.CrossEntropyLoss() (mean reduction; train_lightning.py:77-79,142) forward AND backward:
 * loss[0] = mean_r(logsumexp(x_r) - x_r[y_r]),  grad = (softmax(x_r) - onehot(y_r)) / n_rows. */
int bliss_xent_mean(const float* logits /* [n_rows, n_cls] */, const int64_t* labels, int32_t n_rows,
                    int32_t n_cls, float* row_loss /* [n_rows] scratch */, float* loss /* [1] */,
                    float* grad /* [n_rows, n_cls] */, void* stream);

/* ---- optimizer step ------------------------------------------------------------------------
 * replaces torch.optim.Adam(params, lr).step() (train_lightning.py:205-216; betas / eps given by the
 * caller, no weight decay, no amsgrad) over flat, 16-byte aligned fp32 buffers: parameters, gradients
 * and both moments of the whole model.  lr and the step count are device scalars (CUDA-graph replay:
 * a scheduler changes lr between replays); the call advances *step_dev by one and, with zero_grad,
 * clears the gradient buffer for the next backward pass. */
/* grad[r, c] += sum_s part[s][r][c], c < cols: ordered reduction of the row-chunk partials of a split-K weight
 * gradient ([n_parts, rows, cols_pad], e.g. from a batched GEMM) straight into the parameter's gradient. */
int bliss_splitk_accumulate(const float* part, int32_t n_parts, int32_t rows, int32_t cols_pad, int32_t cols,
                            float* grad, void* stream);
int bliss_adam_step(float* params, float* grads, float* exp_avg, float* exp_avg_sq, int64_t n,
                    const float* lr_dev, float beta1, float beta2, float eps, int64_t* step_dev,
                    int32_t zero_grad, void* stream);
/* data parallel without an all-reduce call (see bliss_grad_p2p) */
int bliss_grad_push(const float* grads, int64_t n, const bliss_grad_p2p* q, void* stream);
int bliss_adam_step_p2p(float* params, float* grads, float* exp_avg, float* exp_avg_sq, int64_t n,
                        const float* lr_dev, float beta1, float beta2, float eps, int64_t* step_dev,
                        const bliss_grad_p2p* q, int32_t* error /* device int or NULL */, void* stream);

/* ---- measurement hooks (bench.py; not on the product path) -------------------------------------
 * Per-KERNEL CUDA-event timing: while enabled, every launch of the hot-path kernels is bracketed by an event
 * pair on its launch stream.  bliss_profile_enable(on) clears what was recorded; bliss_profile_read waits for
 * the recorded events and returns one entry per kernel name (names '\n'-separated, ms = total milliseconds,
 * calls = launches); returns the number of entries, <0 if a buffer is too small.  Do not enable during
 * stream capture. */
int bliss_profile_enable(int32_t on);
int bliss_profile_read(char* names, int32_t names_cap, float* ms, int32_t* calls, int32_t cap);
/* L2 -> SM gather probe: every warp adds `rows_per_warp` pseudo-random rows (dim floats, dim % 128 == 0) of an
 * L2-resident table into registers with `mlp` rows in flight, and stores one row of sums.  bench.py times
 * it to MEASURE the gather bandwidth of the chip the SpMM is bounded by (its source rows are L2-resident),
 * instead of quoting a figure from another chip's guide. */
int bliss_l2_gather_probe(const float* table, int32_t n_rows, int32_t dim, int32_t rows_per_warp,
                          float* out /* [n_warps, dim] */, int32_t n_warps, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* BLISS_B200_H */
