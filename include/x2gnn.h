/*
 * x2gnn.h -- C ABI of libx2gnn.so, the B200 (sm_100a) implementation of the X2-GNN
 * message-passing hot path.
 *
 * The reference (zfwangDP/X2-GNN) is pure Python and has no FFI; its boundary for this
 * path is the Python API of seven modules.  Each entry point below is what the Python
 * drop-in module of the same reference file binds through ctypes:
 *
 *   x2_dij / x2_bonds_*          replace  atom_graph.py:32-35 (calculate_Dij), :42-45 (gen_bonds_mini)
 *   x2_radius_graph_*            batched form of the same two functions (positions -> edge_index)
 *   x2_triplets_*                replace  edge_graph.py:12-30 (vertex_to_edge_2)
 *   x2_envelope_fwd              replaces envelop.py:16-21 (poly_envelop.forward), :23-32
 *   x2_radial_fwd / _bwd         replace  radial_basis_layer.py:36-40 (RadialBasis.forward) + autograd
 *   x2_sbf_table / x2_sbf_fwd    replace  angular_basis_layer.py:80-93 (F_B_2D.forward)
 *   x2_angular_fwd               replaces angular_basis_layer.py:28-32 (AngularBasisLayer.forward)
 *   x2_envelope_bwd / x2_angular_bwd / x2_sbf_bwd   autograd of those three w.r.t. distances / angles
 *   x2_meta_*                    CSR metadata of the line graph (replaces what PyG's
 *                                propagate/softmax/scatter derive from edge_index,
 *                                sbftransformer_conv.py:109,151)
 *   x2_sbfconv_fwd / _bwd        replace  sbftransformer_conv.py:93-162 (forward + message + PyG
 *                                propagate/softmax/aggregate) and its autograd backward
 *
 * Conventions
 *   - every function returns 0 on success or a negative X2_E* code; x2_last_error() gives a
 *     thread-local message.  Nothing throws.
 *   - all data pointers are DEVICE pointers to contiguous row-major buffers owned by the caller;
 *     the library never allocates or frees caller-visible memory and never synchronises the
 *     stream (exception: none).  Scratch space is passed in (`ws`, sized by *_workspace_bytes).
 *   - `stream` is a cudaStream_t passed as void*.  Calls are re-entrant across streams.
 *   - floating tensors are fp32, index tensors of the reference API are int64, internal CSR
 *     metadata is int32 (so E, T < 2^31).
 */
#ifndef X2GNN_H_
#define X2GNN_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define X2_OK 0
#define X2_EINVAL (-1)   /* bad argument / unsupported shape */
#define X2_ECUDA (-2)    /* CUDA runtime error (message has the cudaError string) */
#define X2_EWORKSPACE (-3) /* workspace too small */
#define X2_EDEVICE (-4)  /* device is not sm_100 */

int x2_version(void);
const char* x2_last_error(void);
/* 0 iff `device` exists and is compute capability 10.x (B200). */
int x2_device_check(int device);

/* ---------------------------------------------------------------- instrumentation (bench.py)
 * x2_launch_count: number of kernels this library has launched in this process.
 * x2_timing_*:     optional CUDA-event brackets around the phases of x2_sbfconv_fwd/_bwd, recorded
 *                  on the caller's stream.  read() synchronises on the recorded events, ADDS the
 *                  elapsed milliseconds / call counts per phase into ms[]/calls[] (n entries) and
 *                  clears the recording.  Phase ids: X2_PHASE_*. */
int64_t x2_launch_count(void);
int x2_timing_enable(int on);
int x2_timing_read(double* ms, int64_t* calls, int n);
const char* x2_timing_phase_name(int phase);
#define X2_PHASE_NODE_PROJ 0   /* rbf filter + Q|K|V|skip GEMM */
#define X2_PHASE_TROW_PROJ 1   /* lin_edge / lin_sbf on T rows */
#define X2_PHASE_ATTN_FWD 2    /* segmented attention forward */
#define X2_PHASE_ATTN_BWD_TGT 3
#define X2_PHASE_ATTN_BWD_SRC 4
#define X2_PHASE_TROW_DGRAD 5  /* d edge_attr, d sbf */
#define X2_PHASE_TROW_WGRAD 6  /* dW_edge, dW_sbf, db_sbf */
#define X2_PHASE_NODE_BWD 7    /* node-level wgrad/dgrad, filter backward */
#define X2_NUM_PHASES 8

/* ---------------------------------------------------------------- radius graph
 * atom_graph.py:32-35.  dij[n,n] = relu(sqrt(|a|^2 + |b|^2 - 2 a.b)) (Gram form, fp32). */
int x2_dij(const float* pos, int64_t n, float* dij, void* stream);
/* atom_graph.py:42-45.  adj = (dij < cutoff) & (dij != 0).  rowptr[n+1] int32: exclusive scan of
 * the per-row counts (E = rowptr[n], read back by the caller).  ws >= x2_scan_workspace_bytes(n). */
int x2_bonds_count(const float* dij, int64_t n, float cutoff, int32_t* rowptr, void* ws,
                   size_t ws_bytes, void* stream);
/* edge_index[2,E] int64, lexicographic (i, then j) -- the order np.argwhere produces. */
int x2_bonds_fill(const float* dij, int64_t n, float cutoff, const int32_t* rowptr,
                  int64_t* edge_index, int64_t E, void* stream);
/* Batched: atoms of graph g are ptr[g]..ptr[g+1] (ptr int64 [B+1]), batch[n] int64 sorted.  Same
 * arithmetic as x2_dij per pair; edge_index carries global atom ids, sorted by (i, j). */
int x2_radius_graph_count(const float* pos, const int64_t* batch, const int64_t* ptr, int64_t n,
                          float cutoff, int32_t* rowptr, void* ws, size_t ws_bytes, void* stream);
int x2_radius_graph_fill(const float* pos, const int64_t* batch, const int64_t* ptr, int64_t n,
                         float cutoff, const int32_t* rowptr, int64_t* edge_index, int64_t E,
                         void* stream);
size_t x2_scan_workspace_bytes(int64_t n);

/* ---------------------------------------------------------------- triplets (edge_graph.py:12-30)
 * For each bond e=(i->j) in order, for each bond f=(j->k), k ascending, k != i:
 *   triplets_index[0,t]=f, [1,t]=e, edge_j[t]=j, edge_i[t]=i, edge_k[t]=k.          (int64)
 * count: builds the per-atom out-bond CSR in `ws`, writes rowptr[E+1] (int32, by target bond;
 *        T = rowptr[E]) and flags[0]=1 if edge_index was lexicographically sorted,
 *        flags[1]=number of out-of-range atom ids.
 * fill : needs the same `ws` contents. */
size_t x2_triplets_workspace_bytes(int64_t E, int64_t N);
int x2_triplets_count(const int64_t* edge_index, int64_t E, int64_t N, int32_t* rowptr,
                      int32_t* flags, void* ws, size_t ws_bytes, void* stream);
int x2_triplets_fill(const int64_t* edge_index, int64_t E, int64_t N, const int32_t* rowptr,
                     int64_t T, int64_t* triplets_index, int64_t* edge_j, int64_t* edge_i,
                     int64_t* edge_k, const void* ws, size_t ws_bytes, void* stream);

/* ---------------------------------------------------------------- line-graph CSR metadata
 * edge_index[2,T] int64 (row 0 = source line-node, row 1 = target line-node), E line-nodes.
 * Outputs (int32): src[T], tgt[T]; rowptr_tgt[E+1] + order_tgt[T] (triplet ids grouped by target,
 * ascending inside a group; identity when edge_index[1] is already sorted); rowptr_src[E+1] +
 * order_src[T] (grouped by source, ascending).  flags is int32 [4]: flags[0]=1 if target-sorted,
 * flags[1]=number of out-of-range ids, flags[2]=length of the longest target segment, flags[3]=0. */
size_t x2_meta_workspace_bytes(int64_t T, int64_t E);
int x2_meta_build(const int64_t* edge_index, int64_t T, int64_t E, int32_t* src, int32_t* tgt,
                  int32_t* rowptr_tgt, int32_t* order_tgt, int32_t* rowptr_src,
                  int32_t* order_src, int32_t* flags, void* ws, size_t ws_bytes, void* stream);

/* Work items of the fused tile kernels (csrc/tile_attn.cuh): every target segment of the target-sorted
 * triplet list is cut, from its own start, into ITEMS of at most X2_ITEM_ROWS rows -- the unit of work of an
 * attention warp; X2_UNIT_ITEMS consecutive items form the contiguous row range a CTA streams at a time.
 * x2_items_build fills itemptr[E + 1] (items of target e are itemptr[e] .. itemptr[e+1]) and items[2 * n]
 * int32, n <= x2_items_bound(T, E): (first triplet, (target << 4) | (rows - 1)) in row order, followed by the
 * sentinel (T, 0).  Needs E < 2^27.  ws >= x2_items_workspace_bytes(E). */
#define X2_ITEM_ROWS 8
#define X2_UNIT_ITEMS 15
int64_t x2_items_bound(int64_t T, int64_t E);
size_t x2_items_workspace_bytes(int64_t E);
int x2_items_build(const int32_t* rowptr_tgt, int64_t E, int64_t T, int32_t* itemptr, int32_t* items, void* ws,
                   size_t ws_bytes, void* stream);

/* Closed blocks of a target-sorted line graph (csrc/blk_attn.cuh; SURVEY.md section 7 "structural facts": for
 * the line graph built by edge_graph.py:12-30 a block is "the bonds leaving atom j" with the bonds entering j
 * as its targets).  A block is a contiguous range of line-nodes [blk_sptr[b], blk_sptr[b+1]) such that every
 * target segment draws all of its sources from one block; blk_tptr[b] .. blk_tptr[b+1] indexes blk_tord, the
 * target ids of block b in ascending order.  Every line-node is the source side of exactly one block and the
 * target side of exactly one block.  Nothing is assumed about edge_index beyond target-sortedness: the
 * closure is VERIFIED, flags[0] = 1 iff it holds (otherwise the block kernels must not be used).
 * flags int32 [6]: ok, number of blocks, most triplets in a block, most sources in a block, most targets in a
 * block, 0.  blk_sptr / blk_tptr: int32 [E + 1]; blk_tord: int32 [E]; blk_tpos: int32 [E], position of a target
 * inside its block (blk_tord[blk_tptr[b] + blk_tpos[e]] == e).  ws >= x2_blocks_workspace_bytes(T, E). */
size_t x2_blocks_workspace_bytes(int64_t T, int64_t E);
int x2_blocks_build(const int32_t* src, const int32_t* tgt, const int32_t* rowptr_tgt, int64_t T, int64_t E,
                    int32_t* blk_sptr, int32_t* blk_tptr, int32_t* blk_tord, int32_t* blk_tpos, int32_t* flags,
                    void* ws, size_t ws_bytes, void* stream);

/* ---------------------------------------------------------------- basis expansions
 * envelop.py:16-21: out = 1/x + a x^(p-1) + b x^p + c x^(p+1), x = d * inv_cutoff. */
int x2_envelope_fwd(const float* d, int64_t n, float inv_cutoff, int32_t p, float a, float b,
                    float c, float* out, void* stream);
/* radial_basis_layer.py:36-40: out[e,r] = sin(freq[r] * d[e] * inv_cutoff) * (env ? env[e] : 1).
 * (`env` fuses xgnn.py:69; pass NULL for the plain module.) */
int x2_radial_fwd(const float* d, const float* freq, const float* env, int64_t n, int32_t R,
                  float inv_cutoff, float* out, void* stream);
/* grad_freq[R] (+ optional grad_d[n]) of the above given grad_out[n,R].  Deterministic two-stage
 * reduction; ws >= x2_radial_bwd_workspace_bytes(n, R). */
size_t x2_radial_bwd_workspace_bytes(int64_t n, int32_t R);
int x2_radial_bwd(const float* d, const float* freq, const float* env, const float* grad_out,
                  int64_t n, int32_t R, float inv_cutoff, float* grad_freq, float* grad_d,
                  void* ws, size_t ws_bytes, void* stream);
/* angular_basis_layer.py:81-86: table[e, l*R+n] = env(d_e) * norm[l,n] * j_l(zeros[l,n]*d_e/cutoff)
 * with env = poly envelope (cutoff env_cutoff, exponent p-1).  j_l evaluated in fp64 (series /
 * upward recurrence), rounded once to fp32.  zeros/norm: fp32 [L,R] (basis_func.py:14-29,55-60). */
int x2_sbf_table(const float* d, int64_t n, int32_t L, int32_t R, const float* zeros,
                 const float* norm, float cutoff, float env_cutoff, int32_t p, float a, float b,
                 float c, float* table, void* stream);
/* angular_basis_layer.py:87-93: out[t, l*R+n] = table[idx[t], l*R+n] * Y_l0(angles[t]). */
int x2_sbf_fwd(const float* table, const float* angles, const int64_t* idx, int64_t T, int64_t E,
               int32_t L, int32_t R, float* out, void* stream);
/* angular_basis_layer.py:28-32: out[t,l] = Y_l0(angles[t]). */
int x2_angular_fwd(const float* angles, int64_t T, int32_t L, float* out, void* stream);
/* Geometry of the line graph as the caller computes it before the expansions (xgnn.py:46,60-66):
 *   d[e]   = | pos[a0[e]] - pos[a1[e]] |
 *   ang[t] = atan2(| ji x jk |, <ji, jk>),  ji = pos[ai[t]] - pos[aj[t]],  jk = pos[ak[t]] - pos[aj[t]]
 * pos [N,3] fp32, indices int64 (vertex_to_edge_2's atom ids), fp32 arithmetic in the reference's operation order.
 * Two launches instead of ~15 gathers / elementwise passes / reductions over [T,3] intermediates. */
int x2_bond_lengths(const float* pos, const int64_t* a0, const int64_t* a1, int64_t E, float* d, void* stream);
int x2_triplet_angles(const float* pos, const int64_t* ai, const int64_t* aj, const int64_t* ak, int64_t T, float* ang,
                      void* stream);
/* Gradients of the three expansions w.r.t. the geometry.  The reference's bases are torch expressions
 * (envelop.py:16-21, angular_basis_layer.py:28-32,80-93), so its autograd differentiates them w.r.t. distances
 * and angles (e.g. forces = -dE/dpos); these are the analytic derivatives of the same formulas, E-scale parts
 * in fp64 like the forward, deterministic (no atomics).
 *   x2_envelope_bwd : grad_d[i] = grad_out[i] u'(d_i inv_cutoff) inv_cutoff
 *   x2_angular_bwd  : grad_angles[t] = sum_l grad_out[t,l] dY_l0/dtheta(angles[t])
 *   x2_sbf_bwd      : grad_angles[T] and / or grad_d[E] (either may be NULL) of x2_sbf_fwd(x2_sbf_table(d), ..)
 *                     given grad_out[T, L R]; grad_d needs the triplets grouped by idx: order[T] (stable sort of
 *                     idx), rowptr[E+1].  L R <= 256. */
int x2_envelope_bwd(const float* d, const float* grad_out, int64_t n, float inv_cutoff, int32_t p, float a, float b,
                    float c, float* grad_d, void* stream);
int x2_angular_bwd(const float* angles, const float* grad_out, int64_t T, int32_t L, float* grad_angles, void* stream);
int x2_sbf_bwd(const float* d, const float* table, const float* angles, const int64_t* idx, const int64_t* order,
               const int64_t* rowptr, const float* grad_out, int64_t T, int64_t E, int32_t L, int32_t R,
               const float* zeros, const float* norm, float cutoff, float env_cutoff, int32_t p, float a, float b,
               float c, float* grad_d, float* grad_angles, void* stream);

/* ---------------------------------------------------------------- tensor-core Linear building blocks
 * tcgen05 (kind::tf32) GEMMs in 3xTF32 split precision (fp32-accurate), used inside x2_sbfconv_* for
 * the torch.nn.Linear layers of the reference (sbftransformer_conv.py:99,105-107,121,144,148) and
 * exported so the parity tests can exercise them directly.
 *   x2_tc_gemm : C[M,N] (+)= A[M,K] . B(K,N) + bias[N],  B(k,n) = W[k*sbk + n*sbn], N <= 128.
 *                (forward y = x W^T: sbk=1, sbn=ldw;  dgrad dx = dy W: sbk=ldw, sbn=1)
 *   x2_tc_wgrad: dW[128,N] = Y[rows,128]^T . X[rows,N],  db[128] = column sums of Y (db may be NULL),
 *                N <= 128; deterministic split over rows. */
size_t x2_tc_gemm_workspace_bytes(int32_t K, int32_t N);
int x2_tc_gemm(const float* A, int64_t lda, int64_t M, int32_t K, const float* W, int64_t sbk, int64_t sbn,
               int32_t N, const float* bias, float* C, int64_t ldc, int32_t beta, void* ws, size_t ws_bytes,
               void* stream);
size_t x2_tc_wgrad_workspace_bytes(int64_t rows, int32_t N);
int x2_tc_wgrad(const float* Y, int64_t ldy, const float* X, int64_t ldx, int64_t rows, int32_t N, float* dW,
                int64_t lddw, float* db, void* ws, size_t ws_bytes, void* stream);
/* Several weight gradients over the same `rows` and N in as few launches as possible (16 problems per k_tc_wgrad
 * launch, the CTAs dealt over the problems).  A training step defers the weight gradients of its torch.nn.Linear
 * layers (residual_layer.py, readout.py, model.py:16-20 around the conv layers) to the end of the backward and
 * flushes them through this entry point: nothing downstream of a Linear's backward needs dW before the optimizer.
 * Same workspace bound as x2_tc_wgrad. */
typedef struct {
  const float* Y; int64_t ldy;     /* grad of the Linear's output block [rows, 128] */
  const float* X; int64_t ldx;     /* the Linear's input block [rows, N] */
  float* dW; int64_t lddw;         /* [128, N] block of weight.grad */
  float* db;                       /* [128] block of bias.grad or NULL */
} x2_wgrad_job;
int x2_tc_wgrad_batch(const x2_wgrad_job* jobs, int32_t njobs, int64_t rows, int32_t N, void* ws, size_t ws_bytes,
                      void* stream);

/* ---------------------------------------------------------------- SBFTransformerConv
 * Shapes: x[E,D] rbf[E,R] sbf[T,S] edge_attr[T,A] (A=0 => no lin_edge); D = H*C.
 * Weights use the torch.nn.Linear layout [out,in] of the reference state_dict. */
typedef struct {
  int64_t E, T;
  int32_t D, H, C, S, R, A;
  int32_t fuse_skip;   /* 1: out = attn + lin_skip(x) (concat, root_weight, no beta) */
  int32_t mode;        /* X2_MODE_* */
  float dropout_p;     /* attention dropout (training); 0 disables */
  int32_t tgt_sorted;  /* 1 iff edge_index[1] is non-decreasing (order_tgt is the identity): enables the fused
                          tile kernels; x2_meta_build reports it in flags[0] */
  uint64_t seed;       /* dropout RNG seed */
  /* inputs */
  const float *x, *rbf, *sbf, *edge_attr;
  /* line-graph metadata (x2_meta_build) */
  const int32_t *src, *tgt, *rowptr_tgt, *order_tgt, *rowptr_src, *order_src;
  /* parameters */
  const float *w_rbf;              /* [D,R] */
  const float *w_q, *b_q;          /* [D,D], [D] */
  const float *w_k, *b_k;
  const float *w_v, *b_v;
  const float *w_edge;             /* [D,A] or NULL */
  const float *w_sbf, *b_sbf;      /* [D,S], [D] */
  const float *w_skip, *b_skip;    /* [D,D], [D] (b_skip may be NULL); used iff fuse_skip */
  /* Optional segment-constant edge features (opt-in fast path for callers whose edge_attr row is the
   * same for every triplet of a target line-node, as in xgnn.py:57-58 where it is a function of the
   * central atom only).  ea_index == NULL: edge_attr is [T,A] as in the reference.  ea_index != NULL:
   * edge_attr is a table [ea_rows, A]; every triplet of target e uses row ea_index[e] (int32 [E]);
   * ea_rowptr[ea_rows+1] / ea_order[E] group the targets by table row, ascending inside a group
   * (x2_meta_build on the [2,E] index [ea_index; ea_index] with ea_rows nodes: its rowptr_src /
   * order_src).  Then saved.ea is [ea_rows, D] and grads.dedge_attr is [ea_rows, A]. */
  int64_t ea_rows;
  const int32_t *ea_index, *ea_rowptr, *ea_order;
  /* Optional work items of the target-sorted triplet list (x2_items_build); NULL => unfused kernels.  With them
   * (and tgt_sorted, D == 128, A in {0, 128} or the edge_attr table, S even <= 64, no dropout, no alpha
   * output) lin_edge, lin_sbf and the segmented attention run as ONE kernel that reads edge_attr / sbf once. */
  const int32_t* items;
  const int32_t* itemptr;
  int64_t items_bound;   /* capacity of `items` in entries (>= number of items + 1): sizes the partial-state scratch */
  /* Optional closed blocks of the line graph (x2_blocks_build with flags[0] == 1); nblk == 0 => generic kernels.
   * With them (tgt_sorted, D == 128, no dropout) the backward runs its by-target and by-source passes in ONE
   * kernel, one CTA per block (csrc/blk_attn.cuh). */
  int64_t nblk;
  int32_t blk_max_src;   /* most sources in a block (flags[3]) */
  int32_t sbf_L, sbf_R;  /* factorised sbf: num_spherical, num_radial (S == sbf_L * sbf_R) */
  int32_t blk_max_trip;  /* most triplets in a block (flags[2]) */
  int32_t blk_max_tgt;   /* most targets in a block (flags[4]) */
  int32_t pad0_;
  const int32_t *blk_sptr, *blk_tptr, *blk_tord, *blk_tpos;
  /* Optional factorised form of sbf (angular_basis_layer.py:80-93):
   * sbf[t, l R + n] == sbf_tab[src(t), l R + n] * Y_l0(angles[t]) with sbf_tab [E, S] the per-bond radial
   * table (x2_sbf_table) and angles [T] -- what x2_sbf_fwd multiplies out.  When both are given together
   * with the blocks (and sbf_L <= 8, sbf_R <= 8, grads.dsbf == NULL), lin_sbf is evaluated inside the
   * attention kernels from a per-source table built in shared memory: `sbf` is not read, saved.sg is not
   * written (both may be NULL) and d(lin_sbf out) never exists. */
  const float *sbf_tab, *angles;
} x2_conv_desc;

#define X2_MODE_FP32 0    /* fp32 SIMT arithmetic everywhere (1e-5 parity) */
#define X2_MODE_TF32X3 1  /* Linear layers on tcgen05 tensor cores in 3xTF32 split precision (fp32-accurate,
                             1e-5 parity); needs D % 128 == 0.  Attention arithmetic stays fp32 SIMT. */
#define X2_MODE_TF32 3    /* reduced precision: as TF32X3 but ONE tf32 pass per product (operands truncated to
                             10 mantissa bits by the tensor core, fp32 accumulation): ~5e-4 relative error on
                             layer outputs and gradients, inside the 2e-2 tolerance class of the bf16-projection
                             mode; the lo-half passes of the producers and two of three MMAs are skipped */
#define X2_MODE_TF32X3_FUSED 2  /* as TF32X3, with lin_edge, lin_sbf and the forward attention as ONE tcgen05 kernel over
                                   the work items of x2_items_build (csrc/tile_attn.cuh) when the caller supplies them,
                                   edge_index is target-sorted, D == 128, A in {0, 128}, S even <= 64, no dropout / alpha
                                   request: edge_attr / sbf are read once and EA / Sg are not re-read by the forward
                                   (0.83 GB less DRAM traffic per layer step on the bench batch).  Opt-in: parity-green and
                                   deterministic, but 1 % slower than the unfused kernels on the bench batch today
                                   (profiles/r2_notes.md); X2GNN_FUSED=1 selects it for X2_MODE_TF32X3 too. */

/* Tensors written by fwd and consumed by bwd (caller-owned, kept alive by autograd). */
typedef struct {
  float* qkvs;   /* [E, 4D]  Q | K | V | skip */
  float* attn;   /* [E, D]   attention output before the skip add */
  float* lse;    /* [E, H]   log-sum-exp of the logits per (target, head) */
  float* ea;     /* [T, D]   lin_edge(edge_attr)  (NULL if A == 0) */
  float* sg;     /* [T, D]   lin_sbf(sbf)  (NULL allowed with the factorised sbf of x2_conv_desc) */
  float* xs;     /* [E, D]   x * lin_rbf(rbf), the filtered source features (sbftransformer_conv.py:100); optional:
                    NULL => fwd keeps it in its workspace and bwd recomputes it (one more E-scale kernel) */
} x2_conv_saved;

typedef struct {
  float *dx, *drbf;            /* [E,D], [E,R] */
  float *dsbf;                 /* [T,S] or NULL (not needed by the reference graph) */
  float *dedge_attr;           /* [T,A] or NULL */
  float *dw_rbf, *dw_q, *db_q, *dw_k, *db_k, *dw_v, *db_v, *dw_edge, *dw_sbf, *db_sbf;
  float *dw_skip, *db_skip;    /* used iff fuse_skip (db_skip may be NULL) */
} x2_conv_grads;

/* Which optional paths a descriptor selects (bit set): X2_PLAN_BLOCKS = the backward runs the block-centric
 * kernel; X2_PLAN_FACTORISED_SBF = lin_sbf is evaluated from sbf_tab / angles inside the attention kernels
 * (forward without alpha output and backward), so saved.sg is not needed.  The caller sizes its buffers by it. */
#define X2_PLAN_BLOCKS 1
#define X2_PLAN_FACTORISED_SBF 2
int x2_sbfconv_plan(const x2_conv_desc* d);
size_t x2_sbfconv_fwd_workspace_bytes(const x2_conv_desc* d);
size_t x2_sbfconv_bwd_workspace_bytes(const x2_conv_desc* d);
/* out[E,D]; alpha[T,H] optional (return_attention_weights), NULL otherwise. */
int x2_sbfconv_fwd(const x2_conv_desc* d, const x2_conv_saved* saved, float* out, float* alpha,
                   void* ws, size_t ws_bytes, void* stream);
int x2_sbfconv_bwd(const x2_conv_desc* d, const x2_conv_saved* saved, const float* grad_out,
                   const x2_conv_grads* g, void* ws, size_t ws_bytes, void* stream);

/* ---------------------------------------------------------------- post-conv block
 * Graph-wise LayerNorm without affine parameters: model.py:24,46 `LayerNorm(in_channels, eps=1e-8,
 * affine=False)(x=out, batch=data.batch)` (torch_geometric 2.1.0: mean and variance over all rows AND
 * channels of a molecule).  The rows of molecule g are rowptr[g] .. rowptr[g+1] of x[rowptr[B], D] (PyG
 * collates graph after graph, so they are contiguous); rowptr is int32 [B+1], D % 4 == 0, buffers 16-byte
 * aligned.  fwd writes y[.,D] and stats[B,2] = (mean, 1/sqrt(var+eps)) per molecule; bwd takes the forward's
 * y and stats and grad_y, writes grad_x.  One CTA per molecule, fixed-order reductions (deterministic). */
int x2_graph_layernorm_fwd(const float* x, const int32_t* rowptr, int64_t B, int32_t D, float eps, float* y,
                           float* stats, void* stream);
int x2_graph_layernorm_bwd(const float* y, const float* grad_y, const int32_t* rowptr, int64_t B, int32_t D,
                           const float* stats, float* grad_x, void* stream);

/* rbf-gated bond -> atom readout sum: readout.py:34-43 (AtomWise.forward)
 *   out[n,:] = sum over bonds e = rowptr[n] .. rowptr[n+1]-1 of (w rbf[e] + b) * x[e]
 * for bonds sorted by their first atom (edge_index is lexicographic, atom_graph.py:42-45): rowptr int32 [N+1],
 * rowptr[N] = E; x[E,D], rbf[E,R], w[D,R] (nn.Linear layout), b[D] or NULL; D in {128, 256}, R <= 16.
 * bwd: grad_out[N,D] -> dx[E,D], drbf[E,R], dw[D,R], db[D] (NULL iff b is NULL).  Warp per atom, no atomics,
 * fixed-order reductions (deterministic). */
int x2_rbf_readout_fwd(const float* x, const float* rbf, const float* w, const float* b, const int32_t* rowptr,
                       int64_t N, int64_t E, int32_t D, int32_t R, float* out, void* stream);
size_t x2_rbf_readout_bwd_workspace_bytes(int64_t N, int64_t E, int32_t D, int32_t R);
int x2_rbf_readout_bwd(const float* x, const float* rbf, const float* w, const float* b, const int32_t* rowptr,
                       const float* grad_out, int64_t N, int64_t E, int32_t D, int32_t R, float* dx, float* drbf,
                       float* dw, float* db, void* ws, size_t ws_bytes, void* stream);

/* ---------------------------------------------------------------- parameter-update tail of a training step
 * (trainer.py:43-48, train_ema.py:45-47) over FLAT fp32 buffers of n elements, two launches:
 *   total = ||grad * grad_scale||_2 ; coef = min(1, max_norm / (total + 1e-6))     (torch clip_grad_norm_; max_norm <= 0: off)
 *   g = grad * grad_scale * coef ; exp_avg, exp_avg_sq, param: torch.optim.Adam (defaults: no weight decay / amsgrad)
 *   ema = ema_decay * ema + (1 - ema_decay) * param                               (ema may be NULL)
 * `step` is a device float holding the number of updates done so far; the call increments it (CUDA-graph safe).
 * grad_scale = 1 / world after a sum all-reduce of the flat gradient.  norm_out (device float, may be NULL) receives
 * the gradient norm before clipping.  Deterministic (fixed-order fp64 reduction).  ws >= x2_optim_workspace_bytes(n). */
size_t x2_optim_workspace_bytes(int64_t n);
int x2_optim_tail(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, float* ema, int64_t n,
                  float grad_scale, float max_norm, float lr, float beta1, float beta2, float eps, float ema_decay,
                  float* step, float* norm_out, void* ws, size_t ws_bytes, void* stream);

/* ---------------------------------------------------------------- device-side batch collation
 * (the reference collates PyG records of qm9_allprop.py:18 on the host).  The dataset of M
 * molecules lives on the device in CSR form: atoms of molecule m = rows atom_ptr[m] .. atom_ptr[m+1] of z_all [Na] i64 /
 * pos_all [Na,3] f32, its bonds = rows edge_ptr[m] .. edge_ptr[m+1] of feat_all [Ea,F] f32 and (optional) columns of
 * ei_all [2,Ea] i64 with LOCAL atom ids.  ids [B] i64 selects the batch.
 *   x2_collate_sizes: aoff[B+1], eoff[B+1] int32 exclusive offsets (totals in the last entries); flags[0] = ids out of range
 *   x2_collate_fill : z[N], pos[N,3], batch[N], edge_num[B], feat[E,F], edge_index[2,E] (global ids; NULL to skip) */
size_t x2_collate_workspace_bytes(int64_t B);
int x2_collate_sizes(const int64_t* ids, int64_t B, const int64_t* atom_ptr, const int64_t* edge_ptr, int64_t M,
                     int32_t* aoff, int32_t* eoff, int32_t* flags, void* ws, size_t ws_bytes, void* stream);
int x2_collate_fill(const int64_t* ids, int64_t B, int64_t M, const int64_t* atom_ptr, const int64_t* edge_ptr,
                    const int64_t* z_all, const float* pos_all, const float* feat_all, int32_t F,
                    const int64_t* ei_all, int64_t Etot_all, const int32_t* aoff, const int32_t* eoff, int64_t* z,
                    float* pos, int64_t* batch, int64_t* edge_num, float* feat, int64_t* edge_index, int64_t E_total,
                    void* stream);

#ifdef __cplusplus
}
#endif
#endif /* X2GNN_H_ */
