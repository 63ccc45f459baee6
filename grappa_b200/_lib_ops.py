"""ctypes mirrors of the operator structs / prototypes in include/grappa_b200.h."""
import ctypes as C

vp = C.c_void_p
i32 = C.c_int32
i64 = C.c_int64
f32 = C.c_float
u64 = C.c_uint64


class GemmArgs(C.Structure):
    _fields_ = [
        ("A", vp), ("B", vp), ("C", vp),
        ("M", i32), ("N", i32), ("K", i32),
        ("lda", i32), ("ldb", i32), ("ldc", i32),
        ("trans_a", i32), ("trans_b", i32),
        ("bias", vp),
        ("act", i32),
        ("mul_elu_out", vp),
        ("ldm", i32),
        ("dropout_p", f32),
        ("dropout_seed", u64),
        ("residual", vp),
        ("ldr", i32),
        ("accumulate", i32),
        ("precision", i32),
        ("workspace", vp),
        ("workspace_bytes", i64),
        ("act_out", vp),
        ("ldact", i32),
        ("dropout_offset", vp),
        ("colsum", vp),
        ("ld_colsum", i32),
        ("max_sms", i32),
    ]


COLSUM_MAX = 96


class ColsumDesc(C.Structure):
    _fields_ = [("partial", vp), ("out", vp), ("n_part", i32), ("stride", i32), ("cols", i32), ("accumulate", i32)]


class ColsumBatch(C.Structure):
    _fields_ = [("n", i32), ("pad_", i32), ("desc", ColsumDesc * COLSUM_MAX)]


COLLATE_MAX_JOBS = 64


class CollateJob(C.Structure):
    _fields_ = [("src", vp), ("dst", vp), ("src_off", vp), ("dst_off", vp), ("add", vp), ("confs", vp), ("csel", vp),
                ("row_words", i32), ("kind", i32), ("n_confs_out", i32), ("n_rows", i32)]


class CollateArgs(C.Structure):
    _fields_ = [("n_jobs", i32), ("B", i32), ("mol", vp), ("job", CollateJob * COLLATE_MAX_JOBS)]


MAX_PEERS = 8           # GB_MAX_PEERS
PEER_MAX_CTAS = 64      # GB_PEER_MAX_CTAS


class IpcHandle(C.Structure):
    _fields_ = [("bytes", C.c_ubyte * 64)]


class PeerAllreduceArgs(C.Structure):
    _fields_ = [("data", vp * MAX_PEERS), ("flags", vp * MAX_PEERS), ("rank", i32), ("world", i32), ("start", i64), ("count", i64),
                ("epoch", vp), ("ctas", i32), ("split", i32)]


class Perms(C.Structure):
    _fields_ = [("n_perm", i32), ("perm", (i32 * 4) * 6)]


class FeaturizeArgs(C.Structure):
    _fields_ = [("n_feats", i32), ("feats", vp * 8), ("width", i32 * 8), ("charge", vp), ("enc_dim", i32)]


class HeadOutArgs(C.Structure):
    _fields_ = [
        ("kind", i32), ("T", i32), ("n_perm", i32), ("n_out", i32), ("n_per", i32), ("gated", i32),
        ("k_mean_over_std", f32), ("k_std", f32), ("k_min", f32),
        ("eq_mean_over_std", f32), ("eq_std", f32), ("eq_min", f32),
        ("eq_std_over_max", f32), ("eq_max", f32),
        ("tk_std", f32 * 6), ("tk_mean", f32 * 6),
        ("cutoff", f32),
        ("stat", vp * 4),
    ]


class HeadStatGrads(C.Structure):
    _fields_ = [("d", vp * 4), ("accumulate", i32)]


class LossArgs(C.Structure):
    _fields_ = [
        ("energy", vp), ("energy_ref", vp), ("grad", vp), ("grad_ref", vp), ("atom_off", vp),
        ("k_proper", vp), ("k_improper", vp), ("proper_off", vp), ("improper_off", vp),
        ("B", i32), ("C", i32), ("n_per_p", i32), ("n_per_i", i32),
        ("w_energy", f32), ("w_grad", f32), ("w_proper", f32), ("w_improper", f32),
        ("loss", vp), ("mol_loss", vp), ("g_energy", vp), ("g_grad", vp), ("g_k_proper", vp), ("g_k_improper", vp),
        ("grad_scale", vp), ("extra_mol_loss", vp), ("n_valid", vp),
    ]


class ParamLossArgs(C.Structure):
    _fields_ = [
        ("n_terms", i32), ("B", i32),
        ("pred", vp * 5), ("ref", vp * 5), ("off", vp * 5),
        ("width", i32 * 5), ("ref_width", i32 * 5), ("fac", f32 * 5),
        ("mol_weight", vp), ("mol_loss", vp), ("g_pred", vp * 5), ("grad_scale", vp),
    ]


def declare(lib):
    P = C.POINTER
    lib.grappa_b200_gemm.argtypes = [P(GemmArgs), vp]
    lib.grappa_b200_neighbor_mean.argtypes = [vp, i32, vp, vp, vp, i32, i32, i32, i32, vp]
    lib.grappa_b200_neighbor_mean.restype = C.c_int
    lib.grappa_b200_collate.argtypes = [vp, i32, i64, vp]
    lib.grappa_b200_collate.restype = C.c_int
    lib.grappa_b200_ipc_alloc.argtypes = [i64, P(vp), P(IpcHandle)]
    lib.grappa_b200_ipc_alloc.restype = C.c_int
    lib.grappa_b200_ipc_open.argtypes = [P(IpcHandle), P(vp)]
    lib.grappa_b200_ipc_open.restype = C.c_int
    lib.grappa_b200_ipc_close.argtypes = [vp]
    lib.grappa_b200_ipc_close.restype = C.c_int
    lib.grappa_b200_ipc_free.argtypes = [vp]
    lib.grappa_b200_ipc_free.restype = C.c_int
    lib.grappa_b200_peer_allreduce.argtypes = [P(PeerAllreduceArgs), vp]
    lib.grappa_b200_peer_allreduce.restype = C.c_int
    lib.grappa_b200_pad_rows.argtypes = [vp, i32, i32, i32, vp, i32, vp]
    lib.grappa_b200_pad_rows.restype = C.c_int
    lib.grappa_b200_gemm_can_fuse_colsum.argtypes = [P(GemmArgs)]
    lib.grappa_b200_gemm_can_fuse_colsum.restype = C.c_int
    lib.grappa_b200_gemm_grouped.argtypes = [P(GemmArgs), i32, vp]
    lib.grappa_b200_gemm_grouped.restype = C.c_int
    lib.grappa_b200_layernorm_fwd.argtypes = [vp, vp, vp, vp, vp, vp, i32, i32, f32, vp]
    lib.grappa_b200_layernorm_bwd.argtypes = [vp, vp, vp, vp, vp, vp, i32, i32, vp]
    lib.grappa_b200_col_reduce_workspace.argtypes = [i32, i32]
    lib.grappa_b200_col_reduce_workspace.restype = i64
    lib.grappa_b200_col_reduce.argtypes = [vp, i32, vp, vp, vp, vp, vp, vp, i32, i32, i32, vp]
    lib.grappa_b200_edge_attention_fwd.argtypes = [vp, vp, vp, vp, vp, i32, i32, i32, vp]
    lib.grappa_b200_edge_attention_bwd.argtypes = [vp, vp, vp, vp, vp, vp, vp, vp, i32, i32, i32, vp]
    lib.grappa_b200_tuple_attention_fwd.argtypes = [vp, vp, i32, i32, i32, i32, vp]
    lib.grappa_b200_tuple_attention_bwd.argtypes = [vp, vp, vp, i32, i32, i32, i32, vp]
    lib.grappa_b200_tuple_gather_fwd.argtypes = [vp, i32, vp, vp, vp, i32, i32, i32, i32, vp]
    lib.grappa_b200_tuple_gather_bwd.argtypes = [vp, vp, vp, vp, i32, i32, i32, i32, i32, i32, i32, vp]
    lib.grappa_b200_perm_concat_fwd.argtypes = [vp, vp, P(Perms), i32, i32, i32, vp]
    lib.grappa_b200_perm_concat_bwd.argtypes = [vp, vp, P(Perms), i32, i32, i32, vp]
    lib.grappa_b200_featurize.argtypes = [P(FeaturizeArgs), vp, i32, i32, vp]
    lib.grappa_b200_head_output_fwd.argtypes = [P(HeadOutArgs), vp, vp, vp, vp]
    lib.grappa_b200_head_output_bwd.argtypes = [P(HeadOutArgs), vp, vp, vp, vp, vp]
    lib.grappa_b200_head_output_stats_bwd.argtypes = [P(HeadOutArgs), vp, vp, vp, P(HeadStatGrads), vp]
    lib.grappa_b200_dropout.argtypes = [vp, vp, i64, f32, u64, vp, vp]
    lib.grappa_b200_act_dropout_bwd.argtypes = [vp, vp, vp, i64, f32, u64, vp, vp]
    lib.grappa_b200_axpby.argtypes = [vp, vp, i64, f32, f32, vp]
    lib.grappa_b200_sumsq.argtypes = [vp, i64, vp, vp]
    lib.grappa_b200_sumsq_det.argtypes = [vp, i64, vp, vp, vp]
    lib.grappa_b200_adam_step.argtypes = [vp, vp, vp, vp, i64, f32, f32, f32, f32, i32, vp, f32, f32, vp]
    lib.grappa_b200_adam_step_dev.argtypes = [vp, vp, vp, vp, i64, vp, f32, f32, f32, vp, vp, f32, f32, vp]
    lib.grappa_b200_tick.argtypes = [vp, i32, vp]
    lib.grappa_b200_layernorm_bwd_fused.argtypes = [vp, vp, vp, vp, vp, vp, vp, i32, i32, i32, vp]
    lib.grappa_b200_act_dropout_bwd_fused.argtypes = [vp, vp, vp, vp, i32, i32, i32, f32, u64, vp, vp]
    lib.grappa_b200_finalize_colsums.argtypes = [C.POINTER(ColsumBatch), vp]
    lib.grappa_b200_molwise_loss.argtypes = [P(LossArgs), vp]
    lib.grappa_b200_param_loss.argtypes = [P(ParamLossArgs), vp]
    lib.grappa_b200_param_loss.restype = C.c_int
    for name in ("gemm", "layernorm_fwd", "layernorm_bwd", "col_reduce", "edge_attention_fwd", "edge_attention_bwd",
                 "tuple_attention_fwd", "tuple_attention_bwd", "tuple_gather_fwd", "tuple_gather_bwd",
                 "perm_concat_fwd", "perm_concat_bwd", "featurize", "head_output_fwd", "head_output_bwd", "head_output_stats_bwd", "dropout",
                 "act_dropout_bwd", "axpby", "sumsq", "sumsq_det", "adam_step", "adam_step_dev", "tick", "layernorm_bwd_fused", "act_dropout_bwd_fused", "finalize_colsums", "molwise_loss"):
        getattr(lib, "grappa_b200_" + name).restype = C.c_int
