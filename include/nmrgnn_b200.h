/*
 * nmrgnn_b200 — C ABI of the B200-native (sm_100a) GNN forward path.
 *
 * The reference (ur-whitelab/nmrgnn) has no FFI: its boundary for this path is
 * the Python call `model((atoms, nlist, edges, inv_degree))` on the Keras model
 * returned by `nmrgnn.load_model()` (nmrgnn/library.py:92-103,
 * nmrgnn/model.py:245-274).  This header is the C boundary introduced beneath
 * that call; nmrgnn_b200/_capi.py binds it with ctypes and
 * INTEGRATION.md shows the stub a reference maintainer would add.
 *
 * Conventions
 *   - plain pointers and sizes only; no exceptions cross the boundary;
 *   - every function returns NMRGNN_OK (0) or a negative nmrgnn_status;
 *     nmrgnn_last_error() gives the message for the last failure on a handle;
 *   - `mem` says where the caller's buffers live: NMRGNN_MEM_HOST (the library
 *     copies in/out on its stream; pinned memory makes the copies asynchronous)
 *     or NMRGNN_MEM_DEVICE (pointers are device pointers on the handle's GPU);
 *   - `stream` is a cudaStream_t passed as void*.  NULL = the handle's own
 *     stream and the call returns after the work has completed.  Non-NULL =
 *     work is enqueued on that stream and the call returns immediately
 *     (device buffers only); errors detected on the device (out-of-range
 *     neighbour index) are then reported by nmrgnn_synchronize();
 *   - a handle may be used by one host thread at a time; different handles
 *     (e.g. one per GPU) are independent.  The per-call workspaces belong to the
 *     handle, so its calls execute one after the other: a call on a stream other
 *     than the one of the previous asynchronous call first waits (on the device)
 *     for that call to finish;
 *   - all tensors are dense row-major; float = IEEE binary32; nlist = int32.
 */
#ifndef NMRGNN_B200_H
#define NMRGNN_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define NMRGNN_ABI_VERSION 1

typedef enum nmrgnn_status {
  NMRGNN_OK = 0,
  NMRGNN_ERR_BAD_DIMS = -1,   /* unsupported/inconsistent sizes or null pointers      */
  NMRGNN_ERR_BAD_INDEX = -2,  /* nlist entry outside [0, n_atoms) (TF GatherV2 raises) */
  NMRGNN_ERR_CUDA = -3,       /* CUDA runtime error; message in nmrgnn_last_error      */
  NMRGNN_ERR_OOM = -4,        /* device allocation failed                              */
  NMRGNN_ERR_NO_DEVICE = -5,  /* no sm_100 device / kernels not loadable               */
  NMRGNN_ERR_COMM = -6        /* a peer's gather buffer cannot be mapped (multi-GPU)   */
} nmrgnn_status;

enum { NMRGNN_MEM_HOST = 0, NMRGNN_MEM_DEVICE = 1 };

/* activation ids: keras names used by the reference's hyper-parameters
 * (nmrgnn/model.py:33-36) */
enum { NMRGNN_ACT_LINEAR = 0, NMRGNN_ACT_SOFTPLUS = 1, NMRGNN_ACT_RELU = 2, NMRGNN_ACT_TANH = 3 };

/* Model geometry — the reference's hyper-parameters (nmrgnn/model.py:22-36) plus
 * the element count fixed at GNNModel.build (model.py:236-243). */
typedef struct nmrgnn_dims {
  int32_t num_elem;        /* C: one-hot width of `atoms`                        */
  int32_t atom_features;   /* F: atom_feature_size                               */
  int32_t edge_features;   /* E: edge_feature_size                               */
  int32_t edge_hidden;     /* H: edge_hidden_size == RBF count                   */
  int32_t n_edge_fc;       /* edge_fc_layers (last one linear, H->E)             */
  int32_t n_mp;            /* mp_layers                                          */
  int32_t n_fc;            /* fc_layers (last one F->F/2, no residual)           */
  int32_t mp_activation;   /* NMRGNN_ACT_*                                       */
  int32_t fc_activation;   /* NMRGNN_ACT_*                                       */
  float rbf_low;           /* RBFExpansion low  (nmrgnn/layers.py:126-129)       */
  float rbf_high;          /* RBFExpansion high                                  */
} nmrgnn_dims;

typedef struct nmrgnn_handle nmrgnn_handle;

/* Number of host float arrays nmrgnn_create expects for `dims`:
 * 2*n_edge_fc + 1 + n_mp + 2*n_fc + 2 + 2. */
int nmrgnn_num_weights(const nmrgnn_dims* dims);

/* Create a model on CUDA device `device`.  `weights` are host pointers in this
 * order (shapes row-major):
 *   edge FC i = 0..n_edge_fc-1 : kernel [in,out], bias [out]   (in = H; out = H, last E)
 *   embed kernel [C,F]                                          (Dense, no bias; model.py:241)
 *   MP layer l = 0..n_mp-1     : w [F,F,E]  indexed (l,m,n)     (layers.py:11-18)
 *   FC i = 0..n_fc-1           : kernel [F,out], bias [out]     (out = F, last F/2)
 *   out kernel [F/2,C], out bias [C]                            (model.py:239)
 *   peak_std [C], peak_avg [C]                                  (model.py:222-228,242-243)
 * Replaces: tf.keras.models.load_model(...) reviving the SavedModel
 * (nmrgnn/library.py:101-102). */
int nmrgnn_create(const nmrgnn_dims* dims, const float* const* weights, int n_weights,
                  int device, nmrgnn_handle** out);

void nmrgnn_destroy(nmrgnn_handle* h);

/* peaks[n_atoms] = GNNModel.call((atoms, nlist, edges, inv_degree), training=False)
 * (nmrgnn/model.py:245-274).
 *   atoms      float [n_atoms, C]
 *   nlist      int32 [n_atoms, k]   indices into this call's atoms; padded slots must
 *                                   hold a valid index (0 by convention) and edge 0
 *   edges      float [n_atoms, k]   distances in nm; <= 0 marks a padded slot
 *   inv_degree float [n_atoms]
 * Independent graphs are batched by concatenation (nlist offset per graph).
 * n_atoms == 0 is allowed (no work). */
int nmrgnn_forward(nmrgnn_handle* h, const float* atoms, const int32_t* nlist, const float* edges,
                   const float* inv_degree, int64_t n_atoms, int32_t k, float* peaks,
                   int mem, void* stream);

/* ---- Multi-GPU: batches of independent graphs sharded by whole graphs, one process per GPU ---------------
 * The reference has no multi-device code; graphs never interact (tf.gather indexes inside one graph,
 * nmrgnn/layers.py:33), so the only exchange of the path is "every rank gets every rank's peaks".  It is done
 * over peer memory (NVLink / NVSwitch) by the library's own kernels, without a collective-library call: every
 * rank stores its peaks into its slot of EVERY rank's gather buffer, publishes the call's epoch with a
 * system-scope release, and a wait kernel on the same stream acquires all sources (csrc/peer_gather.cuh).
 *
 *   1. every rank: nmrgnn_comm_local(h, capacity, world, blob)   capacity >= the largest shard (atoms);
 *      allocates this rank's buffers, writes NMRGNN_COMM_HANDLE_BYTES of CUDA IPC handles to `blob`
 *   2. exchange the blobs by any means (torch.distributed / MPI all-gather of bytes; rank order)
 *   3. every rank: nmrgnn_comm_init(h, rank, world, all_blobs)   maps the other ranks' buffers
 *   4. per batch, every rank: nmrgnn_forward_sharded(local shard ..., gathered, mem, stream)
 * `gathered` (may be NULL) receives float [world][capacity_4] (capacity rounded up to a multiple of 4; rank
 * r's peaks are the first n_r entries of row r) in the memory space `mem`; nmrgnn_comm_buffer returns the
 * device-resident copy of the latest call.  A rank with n_local = 0 still calls.  All ranks must make the
 * same sequence of calls.  world = 1 works without peers (steps 2 and IPC are no-ops). */
#define NMRGNN_COMM_HANDLE_BYTES 128
int nmrgnn_comm_local(nmrgnn_handle* h, int64_t capacity, int32_t world, void* ipc_out);
int nmrgnn_comm_init(nmrgnn_handle* h, int32_t rank, int32_t world, const void* all_ipc);
int nmrgnn_forward_sharded(nmrgnn_handle* h, const float* atoms, const int32_t* nlist, const float* edges,
                           const float* inv_degree, int64_t n_local, int32_t k, float* gathered,
                           int mem, void* stream);
const float* nmrgnn_comm_buffer(nmrgnn_handle* h, int64_t* capacity);
void nmrgnn_comm_destroy(nmrgnn_handle* h);

/* Per-block entry points (same `mem`/`stream` rules), one per reference layer,
 * used by the parity tests and by callers that compose their own model. */

/* EdgeFCBlock(RBFExpansion(edges) * mask) * mask   (model.py:251-261, layers.py:137-140)
 *   edges float [n_edges] -> edge_features float [n_edges, E] */
int nmrgnn_edge_features(nmrgnn_handle* h, const float* edges, int64_t n_edges,
                         float* edge_features, int mem, void* stream);

/* embed_layer(atoms): float [n_atoms, C] -> float [n_atoms, F]   (model.py:262) */
int nmrgnn_embed(nmrgnn_handle* h, const float* atoms, int64_t n_atoms, float* nodes,
                 int mem, void* stream);

/* nodes_out = MPLayer_l([nodes_in, nlist, edge_features, inv_degree]) + nodes_in
 * (layers.py:26-46 + the residual of model.py:167); nodes_out must not alias nodes_in. */
int nmrgnn_mp_layer(nmrgnn_handle* h, int32_t layer, const float* nodes_in, const int32_t* nlist,
                    const float* edge_features, const float* inv_degree, int64_t n_atoms, int32_t k,
                    float* nodes_out, int mem, void* stream);

/* peaks = readout(FCBlock(nodes), atoms)   (model.py:191-196, 268-273).
 * fc_nodes (float [n_atoms, F/2]) may be NULL; if not, it receives FCBlock's output. */
int nmrgnn_fc_readout(nmrgnn_handle* h, const float* nodes, const float* atoms, int64_t n_atoms,
                      float* peaks, float* fc_nodes, int mem, void* stream);

/* Device k-nearest-neighbour graph builder: the input producer of the path, i.e.
 * what nmrdata.parse_universe + the inv_degree line compute on the host in the
 * reference (nmrgnn/library.py:111-116, nmrgnn/main.py:239-242).
 *   positions     float [n_atoms, 3] in nm (mem rules as above)
 *   graph_offsets int64 [n_graphs + 1], ALWAYS a host pointer: atoms of graph g are
 *                 rows graph_offsets[g] .. graph_offsets[g+1]-1
 *   k             neighbours per atom (1..32); cutoff_nm <= 0 disables the cutoff
 * Outputs (same `mem` as positions): nlist int32 [n_atoms,k] (batch-global indices,
 * neighbours sorted by distance, self excluded; missing slots = the graph's first
 * atom with distance 0), edges float [n_atoms,k] in nm, inv_degree float [n_atoms]
 * = 1 / #(graph-local index > 0), 0 if none (library.py:115-116 semantics). */
int nmrgnn_knn_graph(nmrgnn_handle* h, const float* positions, const int64_t* graph_offsets,
                     int64_t n_atoms, int64_t n_graphs, int32_t k, float cutoff_nm, int32_t* nlist,
                     float* edges, float* inv_degree, int mem, void* stream);

/* Diagnostic: D[128,128] = A[128,64] @ W[64,128] (host buffers) on the tcgen05 building
 * blocks (K-major 64B-swizzled operands, TMEM accumulators).  mode 0 = 3xTF32 split in one
 * accumulator, mode 1 = single TF32 product, mode 2 = FP16 scaled split with main/correction
 * accumulators (the production scheme, fp32-level accuracy), mode 3 = single FP16 product, mode 4 = mode 3
 * with the A operand in tensor memory, modes 5 / 6 = mode 3 issued by a CTA pair (cta_group::2, M = 256;
 * D = the leader's / the peer's 128 rows, the peer's copy of A rotated by one row). */
int nmrgnn_selftest_gemm(nmrgnn_handle* h, const float* A, const float* W, float* D, int mode);

/* Diagnostic: the round-toward-zero compensation constants of the tensor-core path, in units of
 * 2^-24 (c such that the epilogue multiplies the accumulator by 1 + c * 2^-24): c_ulp[0] = edge-MLP
 * layers (analytic), c_ulp[1 + l] = MP layer l (calibrated against the FFMA kernels when the model is
 * created).  Returns the number of values written (n_mp + 1; 1 if the MP layers are not on tensor
 * cores) or a negative status. */
int nmrgnn_tc_compensation(nmrgnn_handle* h, float* c_ulp, int cap);

/* The edge block RBFExpansion -> mask -> EdgeFCBlock -> mask (model.py:251-261) is a function of one scalar per edge.
 * nmrgnn_create tabulates it once per model in FP64 (nodes every 2^-13 nm up to the distance where every RBF has
 * underflowed, cubic through four nodes per interval, argument reduction exact in fp32) and checks the interpolation
 * error at every interval midpoint; smooth activations only (softplus / tanh / linear).  Returns 1 if the table was
 * accepted (interpolation error <= 2^-27 of the feature scale) and is used while option "edge_table" is 1, 0 if the MLP kernels
 * evaluate every edge; *n_intervals / *rel_error (may be NULL) receive the table size and the measured error. */
int nmrgnn_edge_table_info(nmrgnn_handle* h, int32_t* n_intervals, double* rel_error);

/* Runtime options (value semantics per name):
 *   "tc_compensate" = 0: switch the compensation above off (diagnostics only; default 1);
 *   "tc_min_atoms" = n: calls with fewer than n atoms run on the exact-FP32 kernels (default 1024: a few
 *                     128-row tiles, latency-bound on either path; 0 = always use tensor cores);
 *   "force_ffma" = 1: use the exact-FP32 FFMA kernels even where the tcgen05 path applies;
 *   "edge_table" = 0: evaluate the edge block (RBF -> EdgeFCBlock) with the MLP kernels (tcgen05 / FFMA) for every
 *                  edge instead of the create-time table (default 1 where the table exists, see nmrgnn_edge_table_info);
 *   "knn_warp" = 0: the cell-list search of nmrgnn_knn_graph by eight threads per query atom with per-thread lists in
 *                  shared memory and a merge (round-2 first form) instead of one warp per query atom with the sorted
 *                  list in registers (default 1; identical output);
 *   "knn_cells" = 0: nmrgnn_knn_graph searches every atom of the graph per query (brute force) instead of the cell list
 *                  (default 1; identical output);
 *   "fc_pipe" = 0: node MLP + readout on the round-1 kernel (MMA and epilogue phases alternate on the resident tile; main
 *                  and correction products in separate accumulators) instead of the layer-pipelined kernel
 *                  (kernels_fc_pipe.cuh, default 1: one accumulator per layer, two sets in tensor memory, layer l + 1
 *                  accumulates K-chunk by K-chunk under the epilogue of layer l, the next tile is staged by its own
 *                  warps, the readout is formed from registers);
 *   "fc_pair" = 1: the pipelined node-MLP kernel on CTA pairs (cta_group::2: two tiles per M = 256 instruction, each
 *                  CTA stages half of every W tile); same arithmetic per tile, bit-identical results;
 *   "fc_pos_comp1_x100" = v: slope of the pipelined kernel's position-dependent compensation (48-instruction chains,
 *                  default 50);
 *   "fc_role_counters" = 1 / 2 / 0: diagnostics -- arm / print / disarm per-CTA cycle counters of the pipelined node-MLP
 *                  kernel's warp roles;
 *   "mp_single_acc" = 1: MP layers on the single-accumulator kernel: main and correction products accumulate into ONE
 *                  256-column accumulator, so tensor memory holds two sets and the epilogue of a tile runs under the next
 *                  tile's MMAs (the MMA warp no longer waits for the drain); 7 % faster MP layers, but 144 instead of 48
 *                  truncating accumulation steps per output: max error on config 2 0.83 instead of 0.53 of the tolerance
 *                  (default 0);
 *   "mp_pos_comp1_x100" = v: slope of that kernel's position-dependent compensation (default 55);
 *   "mp_small_tiles" = 0: MP-layer tiles are always 128 atoms (default 1: a call of less than one wave of 128-atom
 *                  tiles -- fewer than 148 x 128 atoms -- runs on tiles of 32 / 64 / 96 atoms spread over more SMs; same
 *                  results, lower latency for single structures);
 *   "mp_l1_prefetch" = 1: the MP kernel's record-loader warp prefetches the tile's own node rows into L1 one feature
 *                  pass ahead of the producers (default 1; a pure hint, no effect on results: -1 % launch time, -4 % with
 *                  "mp_single_acc");
 *   "mp_nsplit" = 1: run the MP layers on column-split CTA pairs (kernels_mp_nsplit.cuh: two CTAs of a cluster share one
 *                  128-atom tile, each owns 128 output columns and gathers 64 rows; double-buffered accumulators) --
 *                  bit-identical results, measured slower than the one-CTA kernel (0.45 vs 0.37 ms); default 0;
 *   "mp_chain_segments" = 1 / 2 / 4 / 8: the MP layers' K loop as that many separately accumulated chains per tile, summed
 *                  in round-to-nearest FP32 by the epilogue warps (default 1 = one 48-instruction chain; re-packs the W'
 *                  images and recalibrates): fewer round-toward-zero steps per chain, lower error, more drains (+13 % /
 *                  +37 % launch time at 2 / 4);
 *   "mp_pos_comp_x100" = v: slope c' = v / 100 x 2^-24 of the MP layers' position-dependent compensation (default 50; 0
 *                  = constant factor only; re-packs the W' images and recalibrates), see DESIGN.md "Accumulation model";
 *   "fc_pos_comp_x100" = v / "edge_pos_comp_x100" = v: the same for the node MLP's 16-instruction chains (default 50)
 *                  and the edge MLP's 8-instruction chains (default 30); 0 = the analytic constant factor instead;
 *   "mp_comp_x10" = c / "mp_comp_delta_x10" = d / "fc_comp_x10" = c: diagnostics (tools/diag_comp_scan.py) -- set every MP
 *                  layer's compensation to c / 10 x 2^-24 (negative: recalibrate), shift it by d / 10 x 2^-24, set the
 *                  node MLP's constant;
 *   "mp_role_counters" = 1 / 2 / 0: diagnostics -- arm per-CTA cycle counters of the warp roles of the next MP-layer
 *                  or edge launch / print their means to stdout / disarm;
 *   "profile"    = 1: nmrgnn_forward records CUDA events (on the launching stream) around its
 *                     stages; read them with nmrgnn_stage_times. */
int nmrgnn_set_option(nmrgnn_handle* h, const char* name, int value);

/* Device time (ms, CUDA events) of each stage of the last profiled nmrgnn_forward, in launch
 * order: edge kernel, embed (+ per-atom max), MP layer 0..n_mp-1, FC+readout.  Waits for that
 * forward to finish.  Returns the number of stages written (n_mp + 3) or a negative status. */
int nmrgnn_stage_times(nmrgnn_handle* h, float* ms, int cap);

/* Wait for everything enqueued through this handle on `stream` (NULL = own stream)
 * and report deferred device-side errors (NMRGNN_ERR_BAD_INDEX). */
int nmrgnn_synchronize(nmrgnn_handle* h, void* stream);

/* Number of CUDA kernels this handle has launched so far (bench.py's gpu_launches). */
int64_t nmrgnn_kernel_launches(const nmrgnn_handle* h);

/* Which compute path the handle selected for its dims, e.g. "tcgen05-fp16x3(edge,mp,fc)" or "ffma". */
const char* nmrgnn_compute_path(const nmrgnn_handle* h);

/* Message for the last error on `h` (h == NULL: last error of a failed create). */
const char* nmrgnn_last_error(const nmrgnn_handle* h);

int nmrgnn_abi_version(void);

#ifdef __cplusplus
}
#endif
#endif /* NMRGNN_B200_H */
