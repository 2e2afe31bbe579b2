/* igmc_b200 — C-ABI of the B200-native IGMC hot path (libigmc_b200.so).
 *
 * The reference (muhanzhang/IGMC) has no FFI: its boundary is the Python API that Main.py
 * star-imports (Main.py:11-15).  The host side in igmc_b200/{util_functions,models,train_eval}.py
 * mirrors that API and calls ONLY the entry points below (via ctypes; INTEGRATION.md shows the
 * binding).  Every pointer is a DEVICE pointer owned by the caller (torch tensors), every call is
 * asynchronous on `stream` (a cudaStream_t passed as void*), returns 0 on success, a negative
 * value for argument errors, or 1000+cudaError for launch failures.  Data-dependent failures
 * (capacity overflow, malformed batch) are reported through the device-side error word `err`
 * (see IGMC_ERR_* in csrc/common.cuh) which the host checks at its next synchronisation point.
 *
 * Each entry point cites the reference code it replaces.
 */
#ifndef IGMC_B200_H
#define IGMC_B200_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define IGMC_MAX_LAYERS 8
#define IGMC_HIDDEN 32      /* latent_dim entries are hard-coded to 32 in Main.py:391 */
#define IGMC_MAX_BASES 4    /* num_bases=4 in Main.py:394 (IGMC's default is 2) */
#define IGMC_LIN1_OUT 128   /* models.py:185 */
#define IGMC_MAX_HOP 3      /* --hop (Main.py:88, default 1); node features are one-hot of width 2h+2 */

/* Device-resident rating matrix.  Replaces SparseRowIndexer / SparseColIndexer
 * (util_functions.py:20-66): flat CSR + CSC, int32 indices, uint8 rating label (= stored value-1). */
typedef struct {
  const int32_t* row_ptr;   /* [num_users+1] */
  const int32_t* col_idx;   /* [nnz] ascending within a row */
  const uint8_t* rating;    /* [nnz] rating label 0..R-1 */
  const int32_t* col_ptr;   /* [num_items+1] */
  const int32_t* row_idx;   /* [nnz] ascending within a column */
  int32_t num_users, num_items;
} igmc_csr_t;

/* Where graph g of a batch takes its (user, item, label) from.  Replaces the `links`/`labels`
 * members of MyDynamicDataset (util_functions.py:119-133) + DataLoader index sampling. */
typedef struct {
  const int64_t* idx;          /* [B] indices into links_*; NULL -> graph g uses entry g */
  const int32_t* links_u;      /* dataset-resident pairs */
  const int32_t* links_v;
  const int32_t* links_label;  /* rating label of the pair (y = class_values[label]); may be NULL */
  const int64_t* pair_id;      /* [B] sampling-stream id per graph; NULL -> idx[g] (or g) */
} igmc_pairs_t;

/* Scratch of igmc_extract_batch (per-graph node lists and counts). */
typedef struct {
  int32_t* nodes_u;  /* [B*cap] global user ids, target first then ascending */
  int32_t* nodes_v;  /* [B*cap] global item ids */
  int32_t* n_u;      /* [B] */
  int32_t* n_v;      /* [B] */
  int32_t* row_cnt;  /* [B*cap] */
  int32_t* m_cnt;    /* [B] undirected edges per graph */
  int32_t* col_cnt;  /* [B*cap] matches per item column */
  int32_t* hop_off;  /* [B*2*(IGMC_MAX_HOP+1)] nodes within distance d per side (users then items); needed for h > 1 */
  int32_t* sync;     /* [B+1] zero-initialised flags of the one-launch h = 1 path (re-armed by the kernel); NULL
                      * selects the generic two-launch path */
} igmc_extract_ws_t;

/* The collated batch in the reference's layout (what construct_pyg_graph + Batch.from_data_list
 * produce, util_functions.py:280-297) plus graph offsets.  Buffers are capacity-sized; the true
 * sizes are counts[0]=N, counts[1]=E (directed). */
typedef struct {
  int32_t node_cap, edge_cap, feat_dim;
  float* x;             /* [node_cap*feat_dim] one-hot of node_label, may be NULL */
  uint8_t* node_label;  /* [node_cap] 0 target user, 1 target item, 2 user, 3 item (h=1) */
  int64_t* batch;       /* [node_cap] graph id per node */
  int32_t* node_gid;    /* [node_cap] global user/item id per node */
  int64_t* edge_index;  /* [2*edge_cap]: row 0 at 0, row 1 at edge_cap */
  int64_t* edge_type;   /* [edge_cap] */
  float* y;             /* [B] */
  int32_t* node_ptr;    /* [B+1] */
  int32_t* edge_ptr;    /* [B+1] directed-edge offsets */
  int32_t* graph_nu;    /* [B] number of user nodes of each graph */
  int32_t* counts;      /* [2] N, E */
  /* optional (all or none): the message-passing adjacency of igmc_adj_t (symmetric form), built in the
   * same pass so that igmc_batch_prepare is not needed for extracted batches; lists sorted by (type, neighbour) */
  int32_t* adj_in_ptr;  /* [node_cap+1] */
  uint32_t* adj_in;     /* [edge_cap] */
  int32_t* adj_eid;     /* [edge_cap] */
  uint64_t* adj_tmp;    /* [edge_cap] scratch */
} igmc_batch_out_t;

/* Enclosing-subgraph extraction + labelling + graph construction + collate for B pairs, h = 1..IGMC_MAX_HOP hops.
 * Replaces MyDynamicDataset.get -> subgraph_extraction_labeling -> construct_pyg_graph
 * (util_functions.py:138-145, 208-297) and PyG's Batch.from_data_list.  `max_nodes_per_hop` < 0
 * means None; `cap` bounds one side's node list over all hops.  `inj_*` (all or none NULL) inject per-graph
 * node lists [B*cap] (test hook: the reference's own random.sample draw; h = 1 only).  `seed_dev` (optional, device) overrides `seed` so that a
 * captured CUDA graph can be replayed with a fresh sampling stream every step.
 * `num_classes` = len(class_values), `max_row_deg` = the longest user row of G (0 = unknown): with W->sync set they
 * let h = 1 batches take the one-launch path (single balanced row scan, look-back offsets; same outputs bit for bit). */
int igmc_extract_batch(const igmc_csr_t* G, const igmc_pairs_t* P, int B, int h, int max_nodes_per_hop,
                       double sample_ratio, uint64_t seed, const uint64_t* seed_dev, int cap,
                       const int32_t* inj_nodes_u, const int32_t* inj_nodes_v,
                       const int32_t* inj_n_u, const int32_t* inj_n_v,
                       const igmc_extract_ws_t* W, const float* class_values, int num_classes, int max_row_deg,
                       const igmc_batch_out_t* O, int* err, void* stream);

/* Device-resident store of pre-extracted subgraphs: the reference's static `MyDataset` keeps
 * `(data, slices)` = every graph's tensors concatenated WITHOUT node offsets plus per-graph boundaries
 * (util_functions.py:92,108-109).  Same content here, in compact types, plus the adjacency. */
typedef struct {
  const int32_t* node_off;    /* [G+1] */
  const int32_t* edge_off;    /* [G+1] directed edges */
  const uint8_t* node_label;  /* [sum n] */
  const int32_t* node_gid;    /* [sum n] */
  const int32_t* edge_src;    /* [sum e] graph-local node ids */
  const int32_t* edge_dst;    /* [sum e] */
  const uint8_t* edge_type;   /* [sum e] */
  const float* y;             /* [G] */
  const int32_t* graph_nu;    /* [G] */
  const int32_t* adj_ptr;     /* [sum n + G] per graph n+1 graph-local list offsets */
  const uint32_t* adj_in;     /* [sum e] */
  const int32_t* adj_eid;     /* [sum e] graph-local directed edge ids */
} igmc_store_t;

/* Mini-batch assembly from the store for graphs idx[0..B): what InMemoryDataset.get + Batch.from_data_list do
 * for the static dataset (train_eval.py:46-51): concat, node-offset edge_index, build `batch`; outputs are the
 * same buffers igmc_extract_batch fills (adjacency included, symmetric form).  One CTA per graph, no host sync. */
int igmc_assemble_batch(const igmc_store_t* S, const int64_t* idx, int B, const igmc_batch_out_t* O, int* err,
                        void* stream);

/* Graph offsets of a foreign (PyG-collated) batch: node_ptr from `batch`, edge_ptr from the graph
 * of each edge's source.  Replaces what Batch.from_data_list knows implicitly. */
int igmc_batch_ptrs(const int64_t* batch, const int64_t* edge_src, int N, int E, int B,
                    int32_t* node_ptr, int32_t* edge_ptr, int* err, void* stream);

/* Message-passing adjacency of a batch: per node the incoming (and outgoing) edge lists sorted by
 * (edge_type, neighbour), entries pack nbr_local | type<<16; *_eid holds the directed edge id used to
 * index dropout draws.  `symmetric` != 0 (extractor output: [u|v ; v|u]) aliases out-lists to in-lists.
 * This is the private structure the fused RGCN kernels consume instead of PyG's per-edge gather
 * (RGCNConv.propagate, third party; call site models.py:201). */
typedef struct {
  int32_t* in_ptr;    /* [node_cap+1] */
  uint32_t* in_adj;   /* [edge_cap] */
  int32_t* in_eid;    /* [edge_cap] */
  int32_t* out_ptr;   /* same three for outgoing edges; ignored when symmetric */
  uint32_t* out_adj;
  int32_t* out_eid;
  uint64_t* tmp;      /* [edge_cap] scratch */
  int32_t symmetric;
} igmc_adj_t;

int igmc_batch_prepare(const int64_t* edge_index, int64_t edge_row_stride, const int64_t* edge_type,
                       const int32_t* node_ptr, const int32_t* edge_ptr, int B, int n_cap,
                       const igmc_adj_t* A, int* err, void* stream);

/* IGMC parameters: one flat fp32 buffer (also the NCCL gradient bucket layout).
 * state_dict names of the reference (models.py:182-185, train_eval.py:168-172):
 *   convs.{l}.att [R,NB] | convs.{l}.basis [NB,in,32] | convs.{l}.root [in,32] | convs.{l}.bias [32]
 *   lin1.weight [128, 2*32*L] | lin1.bias [128] | lin2.weight [1,128] | lin2.bias [1] */
typedef struct {
  int32_t num_layers, num_relations, num_bases, in_dim0;
  int32_t off_att[IGMC_MAX_LAYERS], off_basis[IGMC_MAX_LAYERS], off_root[IGMC_MAX_LAYERS],
      off_bias[IGMC_MAX_LAYERS];
  int32_t off_lin1_w, off_lin1_b, off_lin2_w, off_lin2_b;
  int32_t conv_param_count;  /* params before off_lin1_w */
  int32_t param_count;
  float multiply_by;
  int32_t readout;           /* 0: IGMC target-row readout inside igmc_forward/igmc_backward (models.py:203-215);
                              * 1: none - igmc_forward stops at concat_states, igmc_backward takes S.dstate =
                              *    d loss / d concat_states from an external readout (igmc_sortpool_*) */
  int32_t list_hint;         /* 0 = unknown, else the number of edge-list entries one CTA should be able to stage in
                              * shared memory (largest per-CTA list seen + margin): the cluster plans then give up
                              * staging rows (more, smaller chunks) until the list buffer holds that many - large
                              * subgraphs (ml_100k, 402 nodes) otherwise gather through global memory */
} igmc_model_t;

/* Dropout draws of one step.  edge_keep/hidden_keep (uint8, 1 = keep) inject explicit draws
 * (parity tests); otherwise counter-hash draws keyed by `seed` are used when p > 0. */
typedef struct {
  float adj_dropout;            /* models.py:193-198, 0 disables */
  float hidden_dropout;         /* 0.5 in training (models.py:212), 0 in eval */
  uint64_t seed;
  const uint64_t* seed_dev;     /* optional device word overriding `seed` (CUDA-graph replay) */
  const uint8_t* edge_keep;     /* [E] or NULL */
  const uint8_t* hidden_keep;   /* [B*128] or NULL */
} igmc_dropout_t;

/* Activations kept between forward and backward. */
typedef struct {
  float* states;    /* [node_cap * 32*L]  concat_states (models.py:203) */
  float* zsave;     /* training only, NULL in eval.  cluster plans: [node_cap * R*in0p] 1/deg-scaled relation-space
                     * aggregate of the one-hot input (layer 0 only, in0p = in_dim0 rounded up to 4; layers >= 1 take
                     * their weight gradients from the backward's own aggregate and need nothing saved);
                     * generic plan (cluster 0): [L * node_cap * NB*32] basis-space aggregates of every layer */
  float* inv_deg;   /* [node_cap] 1/max(kept in-degree,1) */
  float* feat;      /* [B * 2*32*L] target-user | target-item rows */
  float* hid;       /* [B * 128] relu(lin1) after dropout scaling */
  float* hid_gscale;/* [B * 128] d hid / d pre-activation */
  float* pred;      /* [B] */
  int32_t* target;  /* [B*2] batch-global index of the target user / item node */
  int32_t node_cap; /* row count of one zsave / dstate layer slab */
  float* dstate;    /* [node_cap * 32*L] d loss / d concat_states written by an external readout; read by
                     * igmc_backward when readout = 1, unused otherwise */
  const float* wprep; /* [L * 2 * 32*((R+1)*32+4)] per-step prepared weights (igmc_prep_weights), cluster plans only */
  long long* prof;    /* optional debug: [grid][64] clock64() stamps of the kernel phases (cluster plans), or NULL */
  int32_t* gate;      /* optional: every CTA of igmc_forward (cluster plans) adds 1 when it starts, see igmc_gate_wait */
} igmc_saved_t;

/* Launch-order gate for work that should run BESIDE the forward kernel without taking SMs from it (the extraction of
 * the next batch on a second stream): a one-warp kernel that returns once `target` more CTAs (= the forward's grid
 * size: all of its one-per-SM clusters are resident) have counted in, or after `timeout_us`.  `gate` = two zero-
 * initialised int32 words (started-CTA count, expected total; never reset).  Enqueue it on the second stream in front
 * of the work to be held back; exactly ONE forward per gate - the one of the same step - must be given `gate` through
 * igmc_saved_t (other forwards pass NULL). */
int igmc_gate_wait(int32_t* gate, int target, int timeout_us, void* stream);

/* Pre-staged edge lists of one batch for the cluster plans (igmc_stage_lists): per (graph, cluster rank) the
 * compacted, kept (after this step's dropout_adj draws, models.py:193-198), pre-swizzled list entries of the rank's
 * own nodes plus the segment tables the gather hands out work from - everything the model kernels otherwise rebuild
 * in their prologues.  The kernels pull an image into shared memory with two bulk (TMA) copies.
 * One image per direction: forward = in-lists, backward = out-lists. */
typedef struct {
  int32_t* tab;       /* [B*cluster][tab_ints] header (entries, segments, staged, own nodes) + segment tables */
  uint32_t* ent;      /* [B*cluster][lcap] staged entries */
  float* inv_deg;     /* [node_cap] 1/max(kept in-degree,1); forward image only (NULL in the backward image) */
  int32_t tab_ints, lcap, chunk, cluster;   /* the consuming kernel's plan, see igmc_stage_plan */
} igmc_stage_t;

/* Shape of the image the forward (backward = 0) / backward (1) kernel of plan `cluster` expects for subgraphs of up
 * to n_cap nodes: fills tab_ints, lcap, chunk, cluster of *img (pointers untouched).  Negative if the plan does not
 * exist. */
int igmc_stage_plan(const igmc_model_t* M, int n_cap, int cluster, int backward, igmc_stage_t* img);

/* Build the images of a prepared batch: `fwd` always, `bwd` when training != 0 (may be NULL otherwise).  The dropout
 * descriptor must be the one the forward / backward of this step will be given (same seed).  Replaces, once per
 * batch and off the critical path, the list staging of igmc_forward / igmc_backward. */
int igmc_stage_lists(const igmc_model_t* M, const int32_t* node_ptr, const int32_t* edge_ptr, const igmc_adj_t* A,
                     int B, int n_cap, const igmc_dropout_t* D, int training, const igmc_stage_t* fwd,
                     const igmc_stage_t* bwd, int* err, void* stream);

/* W_r = sum_b att[r,b] basis[b] (and its transpose) for every layer, once per step, so that the
 * per-subgraph CTAs only copy them into shared memory.  Required before igmc_forward / igmc_backward with
 * cluster > 0 whenever the parameters changed (reference: the same product inside RGCNConv.message and the
 * ARR term, train_eval.py:169-172). */
int igmc_prep_weights(const igmc_model_t* M, const float* params, float* wprep, void* stream);

/* Kernel plan.  `cluster` = 0 selects the generic kernels (csrc/rgcn.cu: one 256-thread CTA per
 * subgraph, any num_relations <= 256); 1/2/4 select the relation-space kernels (csrc/rgcn_rs.cu,
 * num_relations <= 12) with that many CTAs (a thread-block cluster) per subgraph.  Returns the dynamic
 * shared memory in bytes the chosen kernel needs for subgraphs of up to n_cap nodes, or a negative value
 * if the plan is not available (too many relations / does not fit in 227 KB). */
int igmc_model_plan(const igmc_model_t* M, int n_cap, int cluster, int backward);

/* IGMC.forward for a prepared batch (models.py:190-217): 4x (RGCNConv + tanh), concat, target-row
 * readout, lin1/relu/dropout/lin2.  Node features resident in shared memory for all layers.
 * If `y` != NULL also writes dpred[g] = d(mean squared error)/d(lin2 output) for the fused train step
 * (train_eval.py:162), with `loss_scale` = 1/(global number of graphs). */
int igmc_forward(const igmc_model_t* M, const float* params, const uint8_t* node_label,
                 const int32_t* node_ptr, const int32_t* edge_ptr, const igmc_adj_t* A, int B, int n_cap,
                 const igmc_dropout_t* D, int training, const igmc_saved_t* S, const float* y,
                 float loss_scale, float* dpred, float* sqerr, int cluster, const igmc_stage_t* stage, int* err,
                 void* stream);

/* Backward of igmc_forward given dpred [B] (gradient wrt the lin2 output).  Writes partial gradients of
 * the conv parameters, one row per CTA, and the readout factors dhid [B*128]; igmc_grad_reduce turns them into the
 * flat gradient.  Row layout: cluster 0 -> gpart[B][conv_param_count] in parameter layout; cluster > 0 ->
 * gpart[B*cluster][igmc_raw_grad_count()] "raw" rows (per layer dW_r [R][in0p|32][32] | d root | d bias): the
 * (att, basis) chain rule is linear and is applied once to the sum by igmc_grad_reduce.  Must use the same `cluster`
 * as the forward that produced S.  `stage` (optional) = the backward image of igmc_stage_lists.
 * (autograd of models.py:190-217) */
int igmc_backward(const igmc_model_t* M, const float* params, const uint8_t* node_label,
                  const int32_t* node_ptr, const int32_t* edge_ptr, const igmc_adj_t* A, int B, int n_cap,
                  const igmc_dropout_t* D, const igmc_saved_t* S, const float* dpred,
                  float* gpart, float* dhid, int cluster, const igmc_stage_t* stage, int* err, void* stream);

/* igmc_forward (training, with the loss) immediately followed by igmc_backward for the same batch, as ONE launch: a
 * cluster goes from its subgraph's readout straight into its backward (d loss / d pred of a subgraph depends on that
 * subgraph's forward only).  Same arguments, same results bit for bit as the two calls; cluster plans with the IGMC
 * readout only.  What `loss = mse_loss(model(data), y); loss.backward()` spans in train_eval.py:160-175. */
int igmc_forward_backward(const igmc_model_t* M, const float* params, const uint8_t* node_label,
                          const int32_t* node_ptr, const int32_t* edge_ptr, const igmc_adj_t* A, int B, int n_cap,
                          const igmc_dropout_t* D, const igmc_saved_t* S, const float* y, float loss_scale,
                          float* dpred, float* sqerr, float* gpart, float* dhid, int cluster,
                          const igmc_stage_t* stage_fwd, const igmc_stage_t* stage_bwd, int* err, void* stream);

/* floats per raw gradient row of the cluster plans (see igmc_backward) */
int igmc_raw_grad_count(const igmc_model_t* M);

/* grad[p] = sum over gpart rows (conv; `raw_rows` != 0: raw rows of the cluster plans, chain rule
 * d basis[b] = sum_r att[r,b] dW_r, d att[r,b] = <dW_r, basis[b]> applied to the sum) ; lin1/lin2 gradients from
 * (dhid, feat, hid, dpred) ; + ARR * d/dW sum_l sum_r ||W_{r+1}-W_r||^2 (train_eval.py:167-174).  Also writes
 * loss_out[0] = sum_g sqerr[g]*loss_scale + ARR*reg  when loss_out != NULL.  `reg_ws` is a zero-initialised
 * scratch of IGMC_REDUCE_WS_FLOATS floats (partial dot products / regulariser values + int tickets, re-armed by the
 * kernel). */
#define IGMC_REDUCE_WS_FLOATS (IGMC_MAX_LAYERS * 32 * 64 + IGMC_MAX_LAYERS * 32 + 64)
int igmc_grad_reduce(const igmc_model_t* M, const float* params, int B, int gpart_rows, const float* gpart,
                     int raw_rows, const float* dhid, const float* feat, const float* hid, const float* dpred,
                     const float* sqerr, float loss_scale, float arr, float grad_scale,
                     float* grad, float* loss_out, float* reg_ws, void* stream);

/* torch.optim.Adam step (train_eval.py:54,177; lr/weight_decay semantics of torch 1.4 Adam, eps outside
 * the sqrt, no amsgrad) on the flat buffers.  `step_count` points at TWO device int64 words: the step
 * counter (incremented by the kernel) and a zero-initialised completion ticket;
 * grad is multiplied by grad_mul first (1/world after an NCCL sum); `lr_dev` (optional device float)
 * overrides `lr` so LR decay does not invalidate a captured graph.  Optional bookkeeping of the reference's
 * `total_loss += loss.item() * num_graphs` (train_eval.py:176) without a host sync: loss_acc[0] += loss_in[0] *
 * loss_weight. */
int igmc_adam_step(float* params, const float* grad, float* exp_avg, float* exp_avg_sq,
                   int64_t* step_count, int n, float lr, const float* lr_dev, float beta1, float beta2,
                   float eps, float weight_decay, float grad_mul, const float* loss_in, float* loss_acc,
                   float loss_weight, void* stream);

/* ---- data-parallel exchange + fused update (the reference has no DP; world 1 reproduces its step) ----
 * Every rank owns one exchange allocation that all ranks of the node map (CUDA IPC): a double-buffered copy of its
 * gradient plus an array of arrival flags its peers write.  The collective library (NCCL / gloo) only carries the
 * 64-byte handles once at start-up. */
#define IGMC_MAX_RANKS 8
typedef struct {
  float* grad[IGMC_MAX_RANKS];    /* grad[r]: rank r's [2][stride] gradient buffers (r = rank: local memory) */
  int32_t* flag[IGMC_MAX_RANKS];  /* flag[r]: rank r's [IGMC_MAX_RANKS] arrival flags, flag[r][src] = last step src published */
  int64_t* state;                 /* local [4], zero-initialised: exchange step counter, then three int32 tickets */
  int32_t world, rank, stride, pad_;
} igmc_comm_t;

/* cudaMalloc + zero + IPC handle (64 bytes) of an exchange allocation / map a peer's / unmap / free. */
int igmc_comm_alloc(int64_t bytes, void** dev_ptr, unsigned char* handle64);
int igmc_comm_open(const unsigned char* handle64, void** dev_ptr);
int igmc_comm_close(void* peer_ptr);
int igmc_comm_free(void* dev_ptr);

/* igmc_grad_reduce (raw rows) -> all-reduce(SUM) of the flat gradient over the ranks of `comm` -> igmc_adam_step,
 * as ONE kernel: the gradient is assembled into this rank's exchange buffer, the ranks publish / await each other's
 * arrival flags, and every thread sums its elements straight out of the peers' memory (NVLink, fixed rank order ->
 * bit-identical parameters on all ranks) before applying Adam.  Replaces `loss.backward()`'s accumulation +
 * `optimizer.step()` (train_eval.py:175-177) and the ncclAllReduce a DDP port would put between them.
 * `grad_copy` (optional) receives the reduced gradient.  `loss_ring` (optional; `ring_size` a power of two): the
 * step's loss is also stored at loss_ring[step number % ring_size] - pass mapped pinned HOST memory and the host reads
 * every step's loss (`loss.item()`, train_eval.py:176) without a copy launch or a sync.  `wprep` (optional): after
 * the update (second grid barrier) the kernel also rebuilds the prepared weights of igmc_prep_weights from the NEW
 * parameters, so that the next step's forward needs no separate preparation launch.
 * IGMC readout only (readout = 0), cluster plans only. */
int igmc_reduce_update(const igmc_model_t* M, float* params, int B, int gpart_rows, const float* gpart,
                       const float* dhid, const float* feat, const float* hid, const float* dpred,
                       const float* sqerr, float loss_scale, float arr, const igmc_comm_t* comm,
                       float* exp_avg, float* exp_avg_sq, int64_t* step_count, float lr, const float* lr_dev,
                       float beta1, float beta2, float eps, float weight_decay, float grad_mul,
                       float* loss_out, float* loss_acc, float loss_weight, float* reg_ws, float* grad_copy,
                       float* loss_ring, int ring_size, float* wprep, void* stream);

/* ---- SortPooling + 1-D convolution readout of DGCNN_RS (models.py:123-167 over DGCNN.__init__ models.py:65-85) ----
 * Consumes the concat_states a readout=1 igmc_forward produced.  latent_dim = [32,...,32,1] is run by the conv
 * kernels as 32-wide layers whose unused output columns have zero weights, so a states row has `state_stride` =
 * 32*L floats of which the first `width` = sum(latent_dim) are real and the sort key is column width-1.
 * Parameters (same flat bucket, reference state_dict names):
 *   conv1d_params1.weight [c1,1,width] .bias [c1] | conv1d_params2.weight [c2,c1,kw2] .bias [c2] |
 *   lin1.weight [128, dense_dim] .bias [128] | lin2.weight [1,128] .bias [1],  dense_dim = c2 * (k/2 - kw2 + 1). */
typedef struct {
  int32_t k, width, state_stride;
  int32_t c1, c2, kw2;       /* 16, 32, 5 (models.py:75-79) */
  int32_t t1, t2, dense_dim; /* k/2 ; t1 - kw2 + 1 ; c2 * t2 */
  int32_t off_conv1_w, off_conv1_b, off_conv2_w, off_conv2_b, off_lin1_w, off_lin1_b, off_lin2_w, off_lin2_b;
  int32_t param_begin, param_end;   /* the readout parameters' range in the flat bucket */
} igmc_sortpool_t;

typedef struct {
  int32_t* rank;      /* [node_cap] position of every node in its graph's order (last channel descending, ties by index) */
  int32_t* perm;      /* [B*k] batch-global node at position t, -1 = padding (graph smaller than k) */
  float* act1;        /* [B*c1*k]  relu(conv1) */
  float* pool;        /* [B*c1*t1] maxpool */
  float* flat;        /* [B*dense_dim] relu(conv2), channel-major (x.view(len(x), -1), models.py:161) */
  float* hid;         /* [B*128] relu(lin1) after dropout scaling */
  float* hid_gscale;  /* [B*128] */
  float* pred;        /* [B] */
  float* dhid;        /* [B*128] backward */
  float* gpart;       /* [B * (c1*width + c1 + c2*c1*kw2 + c2)] per-graph partial gradients of the two Conv1d */
} igmc_sortpool_saved_t;

/* Dynamic shared memory (bytes) of the readout kernels for graphs of up to n_cap nodes, negative if the
 * description is inconsistent or does not fit in 227 KB. */
int igmc_sortpool_plan(const igmc_sortpool_t* P, int n_cap, int backward);

/* global_sort_pool (PyG 1.4.2) -> Conv1d -> ReLU -> MaxPool1d(2,2) -> Conv1d -> ReLU -> flatten -> lin1 ->
 * ReLU -> Dropout(0.5) -> lin2 -> [:,0] (models.py:155-165), one CTA per graph.  With y != NULL also
 * dpred[g] = d(mean squared error)/d pred and sqerr[g]. */
int igmc_sortpool_forward(const igmc_sortpool_t* P, const float* params, const float* states,
                          const int32_t* node_ptr, int B, int n_cap, const igmc_dropout_t* D, int training,
                          const igmc_sortpool_saved_t* S, const float* y, float loss_scale, float* dpred,
                          float* sqerr, int* err, void* stream);

/* Backward of igmc_sortpool_forward: writes dstate [N * state_stride] (= d loss / d concat_states, zero for nodes
 * that were not pooled and for the padding columns) for igmc_backward, and the readout parameters' gradients
 * (times grad_scale) into grad[param_begin, param_end). */
int igmc_sortpool_backward(const igmc_sortpool_t* P, const float* params, const float* states,
                           const int32_t* node_ptr, int B, int n_cap, const igmc_sortpool_saved_t* S,
                           const float* dpred, float* dstate, float grad_scale, float* grad, int* err,
                           void* stream);

/* Version / build info: returns the compiled SM arch (100) so the host can refuse stale builds. */
int igmc_build_info(void);

#ifdef __cplusplus
}
#endif
#endif
