/* grappa_b200 -- C ABI of the B200-native Grappa hot path (libgrappa_b200.so).
 *
 * Drop-in boundary for the reference's model hot path (hits-mbm-dev/grappa v1.2.1).  The reference
 * is pure Python on torch + DGL and has no FFI of its own; the "operator API" it exposes for this
 * path is the graph-field protocol of `grappa.models`.  Each entry point
 * below replaces the arithmetic of one reference function and is what a maintainer binds (ctypes)
 * from inside the corresponding torch.nn.Module -- see INTEGRATION.md.  Conventions:
 *
 *   - plain pointers and sizes only; all tensors are dense row-major; float = IEEE fp32
 *   - device pointers unless the name says host; nothing is allocated inside the library
 *   - every function enqueues on the `stream` it is given (cudaStream_t passed as void*) and returns
 *     0 or a negative GB_ERR_* code; grappa_b200_last_error() gives the message (thread-local)
 *   - no global state besides a per-process cache of device properties
 */
#ifndef GRAPPA_B200_H
#define GRAPPA_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GB_OK 0
#define GB_ERR_INVALID (-1) /* bad argument / unsupported shape */
#define GB_ERR_CUDA (-2)    /* CUDA runtime / launch failure     */

#define GB_ABI_VERSION 3

const char* grappa_b200_last_error(void);
int grappa_b200_abi_version(void);
/* number of SMs of the current device (negative error code without a GPU) */
int grappa_b200_sm_count(void);
/* number of CUDA kernels this library has launched in the current process (for bench.py's gpu_launches) */
int64_t grappa_b200_launch_count(void);

/* ---------------------------------------------------------------------------------------------
 * Tuple index construction (host code, no GPU needed).
 * Replaces grappa.utils.tuple_indices.get_idx_tuples (reference src/grappa/utils/tuple_indices.py:7-63)
 * and get_torsions (:144-216); orderings are bit-exact.
 * bonds: int64 [n_bonds,2] atom indices.  Outputs are int64.
 * ------------------------------------------------------------------------------------------- */
int grappa_b200_tuples_count(const int64_t* bonds, int64_t n_bonds, int64_t* n_angles, int64_t* n_propers);
int grappa_b200_tuples_build(const int64_t* bonds, int64_t n_bonds, int64_t* bonds_sorted /*[n_bonds,2]*/,
                             int64_t* angles /*[n_angles,3]*/, int64_t* propers /*[n_propers,4]*/);
/* torsions: int64 [n_torsions,4] candidate torsions (any atom order for impropers).  Writes the
 * de-duplicated propers (as given) and, per improper, the 3 cyclic orderings with the central atom
 * at index `central_pos` (reference constants.IMPROPER_CENTRAL_IDX = 2).  out buffers must hold
 * n_torsions*4 and 3*n_torsions*4 int64.  Returns GB_ERR_INVALID for a torsion that is neither. */
int grappa_b200_torsions_classify(const int64_t* bonds, int64_t n_bonds, const int64_t* torsions, int64_t n_torsions,
                                  int central_pos, int64_t* propers_out, int64_t* n_propers_out,
                                  int64_t* impropers_out, int64_t* n_impropers_out);

/* ---------------------------------------------------------------------------------------------
 * Ring-membership atom features without rdkit (host code).  Replaces rdkit_utils.get_ring_encoding (reference
 * utils/rdkit_utils.py:7-24: IsInRing, IsInRingSize(3..8)) for topologies given as a bond list: enc [n_atoms, 7],
 * column 0 = in a ring, column s-2 = on a shortest cycle of s atoms through one of its bonds (3 <= s <= 8).
 * bonds: [n_bonds, 2] atom indices in [0, n_atoms), each bond once.
 * ------------------------------------------------------------------------------------------- */
int grappa_b200_ring_encoding(int64_t n_atoms, const int64_t* bonds, int64_t n_bonds, float* enc);

/* ---------------------------------------------------------------------------------------------
 * Gradient all-reduce over NVLink peer memory (one process per GPU, all GPUs on one NVSwitch node).
 * Replaces the DDP gradient all-reduce the reference gets from pytorch_lightning (reference
 * src/grappa/training/trainrun.py:166-176 builds the pl.Trainer; lightning_model.py:205-230 is the
 * step whose backward it synchronises).  The flat gradient buffer of every rank is allocated with
 * grappa_b200_ipc_alloc and mapped by the other ranks with grappa_b200_ipc_open (handles exchanged by
 * the host, e.g. torch.distributed.all_gather_object); grappa_b200_peer_allreduce then sums one span
 * across the ranks inside ONE kernel: ready barrier, every rank reduces its 1/world slice with direct
 * loads from the peers and stores the sum into every rank's buffer, done barrier.  All ranks must issue
 * the same sequence of calls with the same `ctas`.  Results are bit-identical on all ranks.
 * ------------------------------------------------------------------------------------------- */
#define GB_MAX_PEERS 8
#define GB_PEER_MAX_CTAS 64
typedef struct { unsigned char bytes[64]; } gb_ipc_handle;   /* = cudaIpcMemHandle_t */
/* cudaMalloc (zero-filled) + cudaIpcGetMemHandle */
int grappa_b200_ipc_alloc(int64_t bytes, void** ptr, gb_ipc_handle* handle);
/* map another process's allocation (cudaIpcOpenMemHandle, peer access enabled lazily) */
int grappa_b200_ipc_open(const gb_ipc_handle* handle, void** ptr);
int grappa_b200_ipc_close(void* ptr);
int grappa_b200_ipc_free(void* ptr);
typedef struct {
  float* data[GB_MAX_PEERS];      /* base of every rank's gradient buffer as mapped HERE (own entry = local pointer)       */
  uint32_t* flags[GB_MAX_PEERS];  /* every rank's flag block, GB_MAX_PEERS * GB_PEER_MAX_CTAS words, zero-initialised     */
  int32_t rank, world;
  int64_t start, count;           /* span to reduce, in floats; start must be a multiple of 4                            */
  uint32_t* epoch;                /* LOCAL device memory, GB_PEER_MAX_CTAS + 1 words, zero-initialised: per-CTA launch    */
                                  /* counters; word [GB_PEER_MAX_CTAS] is set to 1 if a barrier wait timed out (~2 s)     */
  int32_t ctas;                   /* grid size, 1..GB_PEER_MAX_CTAS                                                       */
  int32_t split;                  /* 1: the two cross-rank barriers run as one-warp kernels before / after the data      */
                                  /*    kernel (3 launches; waiting for a late rank does not hold `ctas` SMs)             */
} gb_peer_allreduce_args;
int grappa_b200_peer_allreduce(const gb_peer_allreduce_args* a, void* stream);

/* ---------------------------------------------------------------------------------------------
 * MM energy + analytic forces over conformations (kernel K13) and its backward (K14).
 * Replaces internal_coordinates (reference src/grappa/models/internal_coordinates.py:15-125),
 * harmonic_energy / torsion_energy / pool_energy (models/energy.py:8-71) and the
 * torch.autograd.grad call that produces forces (models/energy.py:139).
 * Level order everywhere: 0 = n2 bonds, 1 = n3 angles, 2 = n4 propers, 3 = n4_improper.
 * ------------------------------------------------------------------------------------------- */
typedef struct {
  /* batch geometry */
  const float* xyz;          /* [n_atoms, n_confs, 3] atom-major (reference data/MolData.py:193)     */
  int32_t n_atoms, n_confs, n_mols;
  int32_t max_atoms_per_mol; /* largest molecule of the batch (0 = unknown: global-atomics kernel)   */
  const int32_t* atom_off;   /* [n_mols+1] first atom of each molecule                               */
  /* tuples: global atom indices, per-molecule segment offsets */
  const int32_t* idx[4];     /* [T_l, L_l] with L = 2,3,4,4 ; may be NULL when n_tuples[l] == 0      */
  const int32_t* tup_off[4]; /* [n_mols+1] first tuple of each molecule at level l                   */
  int32_t n_tuples[4];
  /* parameters (fp32). k[0],k[1]: [T] ; k[2],k[3]: [T, n_per] ; eq[0],eq[1]: [T]                    */
  const float* k[4];
  const float* eq[2];
  int32_t n_per[2];          /* periodicities of propers / impropers (<= 6)                          */
  int32_t level_mask;        /* bit l set = level l contributes (Energy(terms=...))                  */
  int32_t offset_torsion;    /* 1: add sum_n |k_n| to every torsion energy (Energy(offset_torsion=True)) */
  /* outputs (any may be NULL) */
  float* energy;             /* [n_mols, n_confs] total bonded energy                                */
  float* term_energy[4];     /* [n_mols, n_confs] per level                                          */
  float* grad;               /* [n_atoms, n_confs, 3] = +dE/dxyz ('gradient', not force)             */
  float* x[4];               /* [T_l, n_confs] internal coordinate (r, theta, phi)                   */
  float* tuple_energy[4];    /* [T_l, n_confs] per-tuple energy                                      */
  /* optional conflict-free schedule (grappa_b200_conflict_free_rounds): enables the atomics-free kernel      */
  const int32_t* sched[4];     /* [n_rounds_l, sched_groups] tuple index or -1                         */
  const int32_t* round_off[4]; /* [n_mols+1] first round of each molecule at level l                   */
  int32_t sched_groups;        /* tuples per round (0 = no schedule)                                   */
  /* per-molecule maxima of the batch (host-computed; 0 = unknown): size the shared-memory copy of one molecule's
   * tuple records in the packed-pair kernel (variant 5) */
  int32_t max_tuples_per_mol[4];
  int32_t max_rounds_per_mol[4];
} gb_energy_args;

/* Conflict-free processing order for K13 (host code).  Within one molecule and level, tuples are packed first-fit
 * into rounds of at most `groups` tuples that share no atom, so `groups` threads can add forces of one round into a
 * shared accumulator with plain read-modify-write (no atomics) and a barrier between rounds.
 *   idx [T, L] int32 global atom indices, tup_off [n_mols+1].  round_off [n_mols+1] receives the first round of every
 *   molecule; sched [capacity_rounds * groups] receives tuple indices (-1 = idle slot).
 * Returns the total number of rounds (call with sched == NULL to size the buffer) or a negative error code. */
int64_t grappa_b200_conflict_free_rounds(const int32_t* idx, const int32_t* tup_off, int32_t n_mols, int32_t L,
                                         int32_t groups, int32_t* round_off, int32_t* sched, int64_t capacity_rounds);

/* Zeroes and fills every non-NULL output.  variant: 0 = auto, 1 = global-atomics kernel (any molecule size),
 * 2 = shared-memory tiled kernel (CTA per molecule x conformation tile, tuple groups, shared atomics),
 * 3 = conformation-per-thread kernel (no atomics; latency-bound, kept for cross-checks),
 * 4 = round-scheduled tiled kernel (needs sched/round_off: no atomics, bit-reproducible),
 * 5 = packed-pair kernel (two conformations per lane on the f32x2 pipe, two schedule rounds per barrier with one force
 *     tile per half-warp; needs sched/round_off with 8 groups and 16-byte aligned index tables; bit-reproducible) --
 *     what `auto` picks when a schedule is given and the molecule's tiles fit in shared memory (<= 177 atoms). */
int grappa_b200_energy_fwd(const gb_energy_args* a, int variant, void* stream);

typedef struct {
  gb_energy_args fwd;        /* same inputs as the forward (outputs ignored)                         */
  const float* g_energy;     /* [n_mols, n_confs] dL/d energy, or NULL                               */
  const float* g_grad;       /* [n_atoms, n_confs, 3] dL/d gradient, or NULL                         */
  float* dk[4];              /* dL/dk  same shapes as k  (overwritten)                               */
  float* deq[2];             /* dL/deq                                                               */
  float* workspace;          /* >= grappa_b200_energy_bwd_workspace() bytes, or NULL                 */
  int64_t workspace_bytes;
} gb_energy_bwd_args;

int64_t grappa_b200_energy_bwd_workspace(const gb_energy_args* a);
int grappa_b200_energy_bwd(const gb_energy_bwd_args* a, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Device-side batch assembly from a dataset resident in HBM.  Replaces, per batch, the reference's host collate
 * (data/GraphDataLoader.py:23-73), set_number_confs (utils/dgl_utils.py:132-171) and batch (utils/dgl_utils.py:11-60)
 * and the host construction of the index tables.  A batch is `n_jobs` concatenations of per-molecule pieces; the
 * argument block itself lives in DEVICE memory (it is part of the one small host->device upload a batch needs:
 * molecule ids, batch offsets, conformation selection, job descriptors).
 *   kind 0: rows of `row_words` 4-byte words copied verbatim          kind 1: int32 rows, non-negative entries + add[b]
 *   kind 2: int64 rows (row_words = 2 * elements) + add[b]            kind 3: conformation field: output row =
 *           n_confs_out x row_words words, conformation c taken from stored conformation csel[b, c]
 *   kind 4: plain copy of n_rows * row_words words (src_off / dst_off unused)
 * src_off [n_dataset_molecules + 1] (int64) = first row (kinds 0-2) or first WORD (kind 3) of every molecule in `src`;
 * dst_off [B + 1] = first output row of every molecule of the batch; mol [B] = dataset molecule ids.
 * ------------------------------------------------------------------------------------------- */
#define GB_COLLATE_MAX_JOBS 64
typedef struct {
  const void* src;
  void* dst;
  const int64_t* src_off;
  const int32_t* dst_off;
  const int32_t* add;          /* [B] or NULL */
  const int32_t* confs;        /* kind 3: [n_dataset_molecules] stored conformations */
  const int32_t* csel;         /* kind 3: [B, n_confs_out] */
  int32_t row_words, kind, n_confs_out, n_rows;
} gb_collate_job;

typedef struct {
  int32_t n_jobs, B;
  const int32_t* mol;
  gb_collate_job job[GB_COLLATE_MAX_JOBS];
} gb_collate_args;

/* dev_args: DEVICE pointer to a gb_collate_args; max_words = largest job (sizes the grid) */
int grappa_b200_collate(const gb_collate_args* dev_args, int32_t n_jobs, int64_t max_words, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Dense linear layer GEMM with fused epilogue.  Replaces every torch.nn.Linear on the path
 * (reference models/graph_attention.py:98-101,125-127,261,267-272; DGL DotGatConv.fc;
 * interaction_parameters.py:148-151; network_utils.py:31-32,105; perm_equiv_transformer.py:231-237)
 * and their autograd backward (dgrad / wgrad).
 *
 *   acc[m,n] = sum_k opA(A)[m,k] * opB(B)[n,k]
 *   v = acc + bias[n]; v = act(v); [act_out[m,n] = v;] v *= elu'(y = mul_elu_out[m,n]);
 *   v *= dropout_mask(seed + *dropout_offset, m*N+n)/(1-p);
 *   v += residual[m,n];  C[m,n] = accumulate ? C[m,n] + v : v
 *
 * trans_a = 0: A is [M,K] row-major (lda >= K);  1: A is stored [K,M] row-major (lda >= M)
 * trans_b = 0: B is [N,K] row-major (nn.Linear weight layout, C = A B^T);  1: B is stored [K,N]
 * precision: 0 = fp32 FFMA (CUDA cores), 1 = TF32 tcgen05 tensor cores (TMA-staged, TMEM accumulator;
 *            needs 16-byte aligned rows, otherwise GB_ERR_INVALID), 2 = auto (TF32 tcgen05 when legal, else FFMA),
 *            3 = bf16x3 tcgen05 when legal, else FFMA: every fp32 operand element is split inside the kernel into
 *            bf16 hi + bf16 lo and the product is hi*hi + lo*hi + hi*lo with fp32 accumulation (16 mantissa bits per
 *            operand -- fp32-class results at tensor-core speed; operands stay plain fp32 arrays in HBM)
 * ------------------------------------------------------------------------------------------- */
typedef struct {
  const float* A;
  const float* B;
  float* C;
  int32_t M, N, K;
  int32_t lda, ldb, ldc;
  int32_t trans_a, trans_b;
  const float* bias;          /* [N] or NULL */
  int32_t act;                /* 0 none, 1 ELU(alpha=1) */
  const float* mul_elu_out;   /* [M,N] (ld = ldm) saved ELU OUTPUT y: multiply by (y > 0 ? 1 : y + 1) */
  int32_t ldm;
  float dropout_p;            /* 0 = off */
  uint64_t dropout_seed;
  const float* residual;      /* [M,N] (ld = ldr) or NULL */
  int32_t ldr;
  int32_t accumulate;
  int32_t precision;
  float* workspace;           /* optional split-K scratch (deterministic partial sums), or NULL */
  int64_t workspace_bytes;
  float* act_out;             /* optional [M,N] (ld = ldact): value after bias+activation, before dropout/residual */
  int32_t ldact;
  const uint64_t* dropout_offset; /* optional DEVICE counter added to dropout_seed at run time: lets a captured CUDA
                                     graph draw a fresh mask on every replay (see grappa_b200_tick) */
  float* colsum;              /* optional [ceil(M/32), ld_colsum]: row g receives the column sums of the STORED values of
                                 rows 32g .. 32g+31 (bias-gradient partials of the layer that produced this GEMM's input,
                                 folded later by grappa_b200_finalize_colsums).  Tensor-core path only and never combined
                                 with split-K: ask grappa_b200_gemm_can_fuse_colsum first */
  int32_t ld_colsum;
  int32_t max_sms;            /* 0 = all.  SMs the persistent tensor-core kernel may occupy: callers that run several
                                 streams side by side (the four writers, the weight-gradient branch) leave part of the
                                 machine to the other streams' kernels (measured optimum 116 of 148) */
} gb_gemm_args;

int grappa_b200_gemm(const gb_gemm_args* a, void* stream);   /* accumulate: 0 = overwrite C, 1 = C += epilogue(acc), 2 = the old C is added to the
                                                                * accumulator BEFORE bias / activation (sum of two Linear layers under one activation) */
/* n independent GEMMs in as few launches as possible: runs of up to 4 tensor-core problems with the same operand
 * layouts, tile configuration and workspace share ONE persistent kernel (work items numbered problem by problem, split-K
 * slices sized so that all problems together fill the SMs once) and one reduce launch; everything else falls back to
 * grappa_b200_gemm per problem.  Results are identical to n separate grappa_b200_gemm calls up to the summation
 * order of split-K slices.  Used for the weight gradients of a layer (reference: torch autograd's per-Linear
 * weight.grad accumulation, e.g. models/network_utils.py:44-54 backward). */
int grappa_b200_gemm_grouped(const gb_gemm_args* list, int32_t n, void* stream);
/* 1 if grappa_b200_gemm would serve `a` with the tensor-core kernel whose epilogue can also emit `colsum`, else 0 */
int grappa_b200_gemm_can_fuse_colsum(const gb_gemm_args* a);

/* ---------------------------------------------------------------------------------------------
 * LayerNorm over the last dimension (eps, affine, biased variance = torch.nn.LayerNorm defaults;
 * reference models/graph_attention.py:258,265, network_utils.py:38,100).
 * ------------------------------------------------------------------------------------------- */
int grappa_b200_layernorm_fwd(const float* x, const float* gamma, const float* beta, float* y, float* mean,
                              float* rstd, int32_t rows, int32_t cols, float eps, void* stream);
/* dx only; parameter gradients come from grappa_b200_col_reduce below */
int grappa_b200_layernorm_bwd(const float* dy, const float* x, const float* mean, const float* rstd,
                              const float* gamma, float* dx, int32_t rows, int32_t cols, void* stream);
/* Column reductions over rows (deterministic):
 *   out_sum[c]  = sum_r dy[r,c]                                  (bias gradients, LayerNorm beta)
 *   out_xhat[c] = sum_r dy[r,c] * (x[r,c] - mean[r]) * rstd[r]   (LayerNorm gamma; skipped if x NULL)
 * One launch: the last block of every column group folds the partial sums in a fixed order.
 * workspace: >= grappa_b200_col_reduce_workspace(rows, cols) bytes whose first 256 bytes (ticket counters) are ZERO
 * when first used; the kernel leaves them zero, so the buffer can be reused by later launches on the same stream.
 * accumulate: add to outputs. */
int64_t grappa_b200_col_reduce_workspace(int32_t rows, int32_t cols);
int grappa_b200_col_reduce(const float* dy, int32_t ld, const float* x, const float* mean, const float* rstd,
                           float* out_sum, float* out_xhat, float* workspace, int32_t rows, int32_t cols,
                           int32_t accumulate, void* stream);

/* Fused backward kernels with DEFERRED column sums.  They write dx and, per CTA, the column sums over the CTA's rows
 * into `partial` ([n_cta, 2, cols] for LayerNorm: set 0 = sum dy (beta), set 1 = sum dy * xhat (gamma);
 * [n_cta, cols] for act_dropout_bwd: sum dx = bias gradient).  grappa_b200_finalize_colsums then folds the partials
 * of up to GB_COLSUM_MAX reductions in ONE launch (fixed order: deterministic).  cols <= 512 (LayerNorm) / <= 8192 (activation-dropout, column slices of 1024). */
int grappa_b200_layernorm_bwd_fused(const float* dy, const float* x, const float* mean, const float* rstd,
                                    const float* gamma, float* dx, float* partial, int32_t n_cta, int32_t rows,
                                    int32_t cols, void* stream);
int grappa_b200_act_dropout_bwd_fused(const float* dy, const float* act_out, float* dx, float* partial, int32_t n_cta,
                                      int32_t rows, int32_t cols, float p, uint64_t seed, const uint64_t* seed_offset,
                                      void* stream);
#define GB_COLSUM_MAX 96
typedef struct {
  const float* partial;   /* first partial row                                   */
  float* out;             /* [cols]                                              */
  int32_t n_part;         /* number of partial rows                              */
  int32_t stride;         /* floats between consecutive partial rows             */
  int32_t cols;
  int32_t accumulate;     /* 1: out += sum                                       */
} gb_colsum_desc;
typedef struct {
  int32_t n;
  int32_t pad_;
  gb_colsum_desc desc[GB_COLSUM_MAX];
} gb_colsum_batch;
int grappa_b200_finalize_colsums(const gb_colsum_batch* batch, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Graph attention over bonded neighbours (DGL DotGatConv body: u_dot_v / sqrt(d) -> edge softmax over
 * the in-edges of each destination -> u_mul_e + sum; call site reference models/graph_attention.py:283).
 * CSR by destination: indptr[n+1], esrc[E]; erev[e] = position of the reverse edge.
 * ft/out/dout/dft: [n_nodes, heads*dim]; alpha, ds: [E, heads].
 * ------------------------------------------------------------------------------------------- */
/* SAGEConv('mean') neighbour aggregation of grappa-1.0's ResidualConvBlock (reference models/graph_attention.py:314-415,
 * dgl.nn.SAGEConv): mode 0: out[v] = mean_{u in N(v)} x[u];  mode 1 (backward; bonded graphs are symmetric):
 * out[u] = sum_{v in N(u)} x[v] / deg(v).  CSR by destination as for edge attention; rows of `width` floats. */
int grappa_b200_neighbor_mean(const float* x, int32_t ldx, const int32_t* indptr, const int32_t* esrc, float* out,
                              int32_t ldo, int32_t n_nodes, int32_t width, int32_t mode, void* stream);
int grappa_b200_edge_attention_fwd(const float* ft, const int32_t* indptr, const int32_t* esrc, float* out,
                                   float* alpha, int32_t n_nodes, int32_t heads, int32_t dim, void* stream);
int grappa_b200_edge_attention_bwd(const float* ft, const float* alpha, const float* dout, const int32_t* indptr,
                                   const int32_t* esrc, const int32_t* erev, float* ds, float* dft,
                                   int32_t n_nodes, int32_t heads, int32_t dim, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Self-attention over the L <= 4 atoms of every tuple (torch.nn.MultiheadAttention core, reference
 * models/network_utils.py:105,122): qkv [L*T, 3E] (row = l*T + t; q | k | v; heads are contiguous
 * head_dim slices), out [L*T, E].  scores = (q / sqrt(head_dim)) k^T, softmax over keys.
 * ------------------------------------------------------------------------------------------- */
int grappa_b200_tuple_attention_fwd(const float* qkv, float* out, int32_t T, int32_t L, int32_t heads,
                                    int32_t head_dim, void* stream);
int grappa_b200_tuple_attention_bwd(const float* qkv, const float* dout, float* dqkv, int32_t T, int32_t L,
                                    int32_t heads, int32_t head_dim, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Tuple gather (RepProjector: atom_feats[idxs].transpose(0,1), reference
 * interaction_parameters.py:173-178) fused with the positional-encoding column
 * (perm_equiv_transformer.py:134-141):  x[l*T+t, 0:F] = p[idx[t,l], 0:F];  x[l*T+t, F:E] = pe[l].
 * Backward is a deterministic segmented sum over the atom -> (tuple, slot) incidence CSR.
 * ------------------------------------------------------------------------------------------- */
int grappa_b200_tuple_gather_fwd(const float* p, int32_t ldp, const int32_t* idx, const float* pe, float* x,
                                 int32_t T, int32_t L, int32_t F, int32_t E, void* stream);
int grappa_b200_tuple_gather_bwd(const float* dx, const int32_t* inv_ptr, const int32_t* inv_ent, float* dp,
                                 int32_t ldp, int32_t n_atoms, int32_t T, int32_t L, int32_t F, int32_t E,
                                 int32_t accumulate, void* stream);

/* Symmetriser input (reference perm_equiv_transformer.py:239-262): for each permutation p of the L
 * positions, s[p*T+t, j*E:(j+1)*E] = x[perm[p][j]*T + t, :].  bwd sums the contributions back. */
typedef struct {
  int32_t n_perm;
  int32_t perm[6][4];
} gb_perms;
int grappa_b200_perm_concat_fwd(const float* x, float* s, const gb_perms* perms, int32_t T, int32_t L, int32_t E,
                                void* stream);
int grappa_b200_perm_concat_bwd(const float* ds, float* dx, const gb_perms* perms, int32_t T, int32_t L, int32_t E,
                                void* stream);

/* ---------------------------------------------------------------------------------------------
 * Input featurisation: concat of per-atom feature tensors + 16-d sinusoidal charge encoding
 * (reference models/graph_attention.py:157-164, 418-444).  feats[i] is [n, width[i]] row-major.
 * out is [n, ld] with ld >= sum(width) + 16; padding columns are zeroed.
 * ------------------------------------------------------------------------------------------- */
typedef struct {
  int32_t n_feats;
  const float* feats[8];
  int32_t width[8];
  const float* charge;   /* [n] partial charges for the encoding, or NULL (charge_encoding=False) */
  int32_t enc_dim;       /* 16 */
} gb_featurize_args;
int grappa_b200_featurize(const gb_featurize_args* a, float* out, int32_t n, int32_t ld, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Output maps of the writers, fused with the sum over the two symmetriser permutations
 * (reference interaction_parameters.py:244-266, 337-362, 519-562; final_layer.py:48-52, 91-97;
 * network_utils.py:144-145).  scores: [n_perm*T, n_out] (row = p*T + t).
 *   kind 0 bond    : eq = ToPositive(c0; eq stats), k = ToPositive(c1; k stats)
 *   kind 1 angle   : eq = max * sigmoid(std_over_max * c0),  k = ToPositive(c1)
 *   kind 2 torsion : gated: k_n = c_n * sigmoid(c_{n+n_per}) * k_std[n]; else c_n * k_std[n] + k_mean[n];
 *                    then zeroed where |k| <= cutoff
 * ------------------------------------------------------------------------------------------- */
typedef struct {
  int32_t kind, T, n_perm, n_out, n_per, gated;
  float k_mean_over_std, k_std, k_min;        /* ToPositive for k  (kinds 0, 1) */
  float eq_mean_over_std, eq_std, eq_min;     /* ToPositive for eq (kind 0)     */
  float eq_std_over_max, eq_max;              /* ToRange for eq    (kind 1)     */
  float tk_std[6], tk_mean[6];                /* torsion statistics (kind 2)    */
  float cutoff;
  /* learnable_statistics=True (reference final_layer.py:37-39,84-85; interaction_parameters.py:465-467): the statistics
   * are parameters, read from DEVICE memory at run time so that a captured step sees every optimizer update.  A non-NULL
   * entry overrides the scalar above.  kind 0: {k mean_over_std, k std, eq mean_over_std, eq std}; kind 1: {k
   * mean_over_std, k std, eq std_over_max, NULL}; kind 2: {k_std[n_per], k_mean[n_per], NULL, NULL}. */
  const float* stat[4];
} gb_head_out_args;
/* Where head_output_stats_bwd writes the gradients of the statistics, slot by slot like gb_head_out_args.stat
 * (NULL = not wanted); accumulate != 0 adds to the existing values. */
typedef struct {
  float* d[4];
  int32_t accumulate;
} gb_head_stat_grads;
int grappa_b200_head_output_fwd(const gb_head_out_args* a, const float* scores, float* k, float* eq, void* stream);
/* dscores [n_perm*T, n_out] from dk, deq (either may be NULL = zero) */
int grappa_b200_head_output_bwd(const gb_head_out_args* a, const float* scores, const float* dk, const float* deq,
                                float* dscores, void* stream);
/* d loss / d statistics from dk, deq (either may be NULL = zero): one CTA, fixed-order reduction over the tuples */
int grappa_b200_head_output_stats_bwd(const gb_head_out_args* a, const float* scores, const float* dk, const float* deq,
                                      const gb_head_stat_grads* out, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Small fused elementwise kernels.
 * ------------------------------------------------------------------------------------------- */
/* y[i] = x[i] * keep(seed + *seed_offset, i) / (1 - p)   (same call regenerates the mask in the backward pass;
 * seed_offset: optional device counter, NULL = 0) */
int grappa_b200_dropout(const float* x, float* y, int64_t n, float p, uint64_t seed, const uint64_t* seed_offset,
                        void* stream);
/* dx[i] = dy[i] * keep(seed, i)/(1-p) * elu'(act_out[i])   (act_out NULL: no activation; p = 0: no mask) */
int grappa_b200_act_dropout_bwd(const float* dy, const float* act_out, float* dx, int64_t n, float p, uint64_t seed,
                                const uint64_t* seed_offset, void* stream);
/* y = a*x + b*y */
int grappa_b200_axpby(const float* x, float* y, int64_t n, float a, float b, void* stream);
/* out[r, c] = c < cols ? in[r, c] : 0 for c < ld_out: zero-padded copy of a matrix whose row pitch is not a 16-byte
 * multiple (GrappaGNN.pre_dense, 85 input features, reference models/graph_attention.py:166) so that its GEMM is
 * TMA-legal */
int grappa_b200_pad_rows(const float* in, int32_t rows, int32_t cols, int32_t ld_in, float* out, int32_t ld_out, void* stream);
/* out[0] += sum x^2  (out must be zeroed by the caller; deterministic two-stage when ws given) */
int grappa_b200_sumsq(const float* x, int64_t n, float* out, void* stream);
/* out[0] = sum x^2, bit-reproducible (fixed grid, partials folded in block order by the last block).
 * workspace: >= 4096 bytes, ZERO on first use (the kernel leaves its ticket counter zero). */
int grappa_b200_sumsq_det(const float* x, int64_t n, float* out, float* workspace, void* stream);
/* Adam (torch.optim.Adam semantics, weight_decay = 0) with the gradient pre-scaled by
 * min(1, clip / (sqrt(*gnorm_sq) + 1e-6)) * grad_scale -- gnorm_sq is a device scalar or NULL. */
int grappa_b200_adam_step(float* p, const float* g, float* m, float* v, int64_t n, float lr, float beta1, float beta2,
                          float eps, int32_t step, const float* gnorm_sq, float clip, float grad_scale, void* stream);
/* Same, with the learning rate and the step count read from DEVICE memory (CUDA-graph replays: the host updates
 * *lr_dev with a copy, grappa_b200_tick advances *step_dev).  *step_dev counts completed steps (0 on the first call). */
int grappa_b200_adam_step_dev(float* p, const float* g, float* m, float* v, int64_t n, const float* lr_dev, float beta1,
                              float beta2, float eps, const uint64_t* step_dev, const float* gnorm_sq, float clip,
                              float grad_scale, void* stream);
/* counters[i] += 1 for i < n  (device-side step / RNG counters advanced inside a captured graph) */
int grappa_b200_tick(uint64_t* counters, int32_t n, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Molecule-wise loss (reference training/loss.py:45-167, terms active in grappa-1.2 training):
 *   loss = mean_b [ w_e * MSE_c(E_b - mean_c E_b, Eref_b - mean_c Eref_b) + w_g * MSE(grad_b, gradref_b)
 *                   + w_p * mean(k_proper_b^2) + w_i * mean(k_improper_b^2) ]
 * and its gradients w.r.t. energy, gradient and the torsion parameters, in one pass.
 * ------------------------------------------------------------------------------------------- */
typedef struct {
  const float* energy;      /* [B,C] */
  const float* energy_ref;  /* [B,C] */
  const float* grad;        /* [N,C,3] */
  const float* grad_ref;    /* [N,C,3] */
  const int32_t* atom_off;  /* [B+1] */
  const float* k_proper;    /* [T4, n_per_p] or NULL */
  const float* k_improper;  /* [T4i, n_per_i] or NULL */
  const int32_t* proper_off;   /* [B+1] */
  const int32_t* improper_off; /* [B+1] */
  int32_t B, C, n_per_p, n_per_i;
  float w_energy, w_grad, w_proper, w_improper;
  float* loss;              /* [1]  (overwritten) */
  float* mol_loss;          /* [B] per-molecule terms (workspace / diagnostics) */
  float* g_energy;          /* [B,C] dloss/denergy   or NULL */
  float* g_grad;            /* [N,C,3]               or NULL */
  float* g_k_proper;        /* like k_proper         or NULL */
  float* g_k_improper;      /* like k_improper       or NULL */
  const float* grad_scale;  /* device scalar multiplied into every g_* output (upstream dL), or NULL = 1 */
  const float* extra_mol_loss; /* [B] or NULL: per-molecule terms computed elsewhere (grappa_b200_param_loss),
                                  added to each molecule's term before the mean over molecules */
  const int32_t* n_valid;      /* [B] or NULL: molecule b's first n_valid[b] conformations are real, the rest are the
                                  padding set_number_confs appends ('is_dummy', reference utils/dgl_utils.py:132-171) and
                                  are ignored exactly as unbatch() / delete_dummy_confs (:63-118) drops them */
} gb_loss_args;
int grappa_b200_molwise_loss(const gb_loss_args* a, void* stream);

/* Classical-parameter term of MolwiseLoss (reference training/loss.py:70-113): per molecule b,
 *   mol_loss[b] = mol_weight[b] * mean over all elements of the concatenated terms of ( fac_i * (p - p_ref) )^2
 * where elements whose reference is NaN count as zero difference (loss.py:101-103) but still count in the mean,
 * and a torsion reference with a different number of periodicities is zero-padded / truncated to the model's
 * (correct_torsion_shape, loss.py:170-182).  Terms in reference order: n2_k, n2_eq, n3_k, n3_eq, n4_k
 * (impropers are excluded, loss.py:91-92).  With g_pred set, also writes
 *   g_pred_i = grad_scale / B * dmol_loss[b]/dp.                                                     */
#define GB_PARAM_LOSS_MAX_TERMS 5
typedef struct {
  int32_t n_terms, B;
  const float* pred[GB_PARAM_LOSS_MAX_TERMS];   /* [T_i, width_i] contiguous */
  const float* ref[GB_PARAM_LOSS_MAX_TERMS];    /* [T_i, ref_width_i] contiguous, NaN = no reference */
  const int32_t* off[GB_PARAM_LOSS_MAX_TERMS];  /* [B+1] tuple offsets of the term's level */
  int32_t width[GB_PARAM_LOSS_MAX_TERMS], ref_width[GB_PARAM_LOSS_MAX_TERMS];
  float fac[GB_PARAM_LOSS_MAX_TERMS];
  const float* mol_weight;                      /* [B] param_weight per molecule */
  float* mol_loss;                              /* [B] (overwritten) */
  float* g_pred[GB_PARAM_LOSS_MAX_TERMS];       /* like pred, or NULL */
  const float* grad_scale;                      /* device scalar or NULL = 1 */
} gb_param_loss_args;
int grappa_b200_param_loss(const gb_param_loss_args* a, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* GRAPPA_B200_H */
