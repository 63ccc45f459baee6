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

#define GB_ABI_VERSION 1

const char* grappa_b200_last_error(void);
int grappa_b200_abi_version(void);
/* number of SMs of the current device (negative error code without a GPU) */
int grappa_b200_sm_count(void);

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
} gb_energy_args;

/* Zeroes and fills every non-NULL output.  variant: 0 = auto, 1 = global-atomics kernel,
 * 2 = shared-memory tiled kernel (one CTA per molecule x conformation tile). */
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

#ifdef __cplusplus
}
#endif
#endif /* GRAPPA_B200_H */
