"""Per-kernel parity through the C ABI against plain torch restatements of the same op (fp64 where
cheap).  fp32 kernels: 1e-5 relative (norm-wise); TF32 tensor-core GEMM: 1e-3."""
import math

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from util import rel_err

pytestmark = pytest.mark.gpu


def R(a, b):
    return rel_err(a.detach().double().cpu().numpy(), b.detach().double().cpu().numpy())


@pytest.fixture(autouse=True)
def _seed():
    torch.manual_seed(0)


@pytest.mark.parametrize("ta,tb", [(False, False), (False, True), (True, False), (True, True)])
@pytest.mark.parametrize("M,N,K", [(1, 1, 1), (37, 85, 19), (128, 128, 64), (300, 511, 256), (1664, 512, 85), (130, 6, 256)])
def test_gemm_fp32_all_layouts(ta, tb, M, N, K):
    from grappa_b200 import ops
    dev = "cuda"
    A = torch.randn((K, M) if ta else (M, K), device=dev)
    B = torch.randn((K, N) if tb else (N, K), device=dev)
    ref = (A.double().T if ta else A.double()) @ (B.double() if tb else B.double().T)
    out = ops.gemm(A, B, trans_a=ta, trans_b=tb, precision=ops.FP32)
    assert R(out, ref) < 1e-5


def test_gemm_fp32_epilogue_and_strided_views():
    from grappa_b200 import ops
    dev = "cuda"
    M, N, K = 203, 511, 256
    X = torch.randn(M, K, device=dev)
    W = torch.randn(N, K, device=dev) / math.sqrt(K)
    b = torch.randn(N, device=dev)
    res = torch.randn(M, 512, device=dev)
    out = torch.zeros(M, 512, device=dev)
    ops.gemm(X, W, bias=b, act=1, residual=res[:, :N], out=out[:, :N], precision=ops.FP32)
    ref = F.elu(X.double() @ W.double().T + b.double()) + res[:, :N].double()
    assert R(out[:, :N], ref) < 1e-5
    assert out[:, N:].abs().max() == 0
    # dgrad with ELU-derivative mask and accumulate
    Y = F.elu(torch.randn(M, K, device=dev))
    dY = torch.randn(M, N, device=dev)
    acc = torch.randn(M, K, device=dev)
    ref = acc.double() + (dY.double() @ W.double()) * torch.where(Y > 0, 1.0, Y.double() + 1.0)
    got = ops.gemm(dY, W, trans_b=True, mul_elu_out=Y, out=acc.clone(), accumulate=True, precision=ops.FP32)
    assert R(got, ref) < 1e-5
    # wgrad (split-K path: long reduction, small output)
    rows = 9000
    dYl = torch.randn(rows, 64, device=dev)
    Xl = torch.randn(rows, 96, device=dev)
    got = ops.gemm(dYl, Xl, trans_a=True, trans_b=True, precision=ops.FP32)
    assert R(got, dYl.double().T @ Xl.double()) < 1e-5


def test_gemm_dropout_mask_is_reproducible_and_unbiased():
    from grappa_b200 import ops
    dev = "cuda"
    X = torch.randn(512, 64, device=dev)
    W = torch.randn(256, 64, device=dev)
    plain = ops.gemm(X, W, precision=ops.FP32)
    d1 = ops.gemm(X, W, dropout_p=0.3, dropout_seed=77, precision=ops.FP32)
    d2 = ops.gemm(X, W, dropout_p=0.3, dropout_seed=77, precision=ops.FP32)
    assert torch.equal(d1, d2)
    keep = d1 != 0
    assert abs(keep.float().mean().item() - 0.7) < 0.01
    assert R(d1[keep], plain[keep] / 0.7) < 1e-6
    # the standalone dropout kernel regenerates the same mask (used by the backward pass)
    m = ops.dropout(torch.ones_like(plain), 0.3, 77)
    assert torch.equal(m != 0, keep)


@pytest.mark.parametrize("rows,cols", [(1, 256), (77, 512), (1000, 1024), (333, 1536), (64, 2048), (5, 128), (9, 64)])
def test_layernorm_fwd_bwd(rows, cols):
    from grappa_b200 import ops
    dev = "cuda"
    x = torch.randn(rows, cols, device=dev) * 2 + 0.3
    g = torch.randn(cols, device=dev)
    b = torch.randn(cols, device=dev)
    y, mean, rstd = ops.layernorm_fwd(x, g, b)
    xd = x.double().requires_grad_(True)
    gd = g.double().requires_grad_(True)
    bd = b.double().requires_grad_(True)
    ref = F.layer_norm(xd, (cols,), gd, bd, 1e-5)
    assert R(y, ref) < 1e-5
    dy = torch.randn(rows, cols, device=dev)
    ref.backward(dy.double())
    dx = ops.layernorm_bwd(dy, x, mean, rstd, g)
    assert R(dx, xd.grad) < 1e-5
    dbeta, dgamma = ops.col_reduce(dy, x=x, mean=mean, rstd=rstd)
    assert R(dbeta, bd.grad) < 1e-5
    assert R(dgamma, gd.grad) < 1e-5


def _random_graph(n_mols=5):
    from grappa_b200 import synthetic
    from grappa_b200.pack import PackedBatch
    g = synthetic.espaloma_mix_batch(seed=4, batch_size=n_mols, n_confs=2).to("cuda")
    return g, PackedBatch(g, device="cuda")


@pytest.mark.parametrize("H,D", [(16, 32), (4, 32), (2, 8), (8, 64)])
def test_edge_attention_fwd_bwd(H, D):
    import grappa_oracle as orc
    from grappa_b200 import ops
    g, pack = _random_graph()
    n = g.num_nodes("n1")
    ft = torch.randn(n, H * D, device="cuda")
    out, alpha = ops.edge_attention_fwd(ft, pack, H)
    src, dst = g.edges()
    ftd = ft.double().cpu().view(n, H, D).requires_grad_(True)
    ref = orc.dot_gat(ftd, src.long().cpu(), dst.long().cpu())
    assert R(out.cpu(), ref.flatten(1)) < 1e-5
    # attention weights sum to one over the in-edges of every destination
    sums = torch.zeros(n, H, device="cuda").index_add(0, torch.repeat_interleave(
        torch.arange(n, device="cuda"), torch.diff(pack["indptr"]).long()), alpha)
    assert (sums - 1).abs().max() < 1e-5
    dout = torch.randn(n, H * D, device="cuda")
    ref.backward(dout.double().cpu().view(n, H, D))
    dft = ops.edge_attention_bwd(ft, alpha, dout, pack, H)
    assert R(dft.cpu(), ftd.grad.flatten(1)) < 1e-5


@pytest.mark.parametrize("L,heads,HD,T", [(2, 8, 64, 50), (3, 8, 64, 33), (4, 8, 64, 129), (4, 4, 32, 7), (3, 2, 16, 1)])
def test_tuple_attention_fwd_bwd(L, heads, HD, T):
    from grappa_b200 import ops
    E = heads * HD
    qkv = torch.randn(L * T, 3 * E, device="cuda")
    out = ops.tuple_attention_fwd(qkv, T, L, heads)
    qd = qkv.double().requires_grad_(True)
    q, k, v = qd.view(L, T, 3 * E).split(E, dim=-1)
    sh = lambda t: t.reshape(L, T, heads, HD).permute(1, 2, 0, 3)
    att = torch.softmax(sh(q) / math.sqrt(HD) @ sh(k).transpose(-1, -2), dim=-1)
    ref = (att @ sh(v)).permute(2, 0, 1, 3).reshape(L * T, E)
    assert R(out, ref) < 1e-5
    dout = torch.randn(L * T, E, device="cuda")
    ref.backward(dout.double())
    dqkv = ops.tuple_attention_bwd(qkv, dout, T, L, heads)
    assert R(dqkv, qd.grad) < 1e-5


@pytest.mark.parametrize("lvl,F,E", [(0, 512, 512), (1, 511, 512), (2, 511, 512), (3, 127, 128)])
def test_tuple_gather_and_perm_concat(lvl, F, E):
    from grappa_b200 import ops
    g, pack = _random_graph(8)
    L = (2, 3, 4, 4)[lvl]
    T = pack.n_tuples[lvl]
    n = pack.n_atoms
    p = torch.randn(n, E, device="cuda")
    pe = torch.tensor([0.0, 1.0, 1.0, 0.0][:L], device="cuda") if F < E else None
    idx = pack[f"idx{lvl}"]
    x = ops.tuple_gather_fwd(p, idx, pe, T, L, F, E)
    ref = p[idx.long()].transpose(0, 1).reshape(L * T, E).clone()
    if F < E:
        ref[:, F:] = pe.repeat_interleave(T)[:, None]
    assert torch.equal(x, ref)
    dx = torch.randn(L * T, E, device="cuda")
    dp = ops.tuple_gather_bwd(dx, pack[f"inv_ptr{lvl}"], pack[f"inv_ent{lvl}"], n, E, T, L, F, E)
    refd = torch.zeros(n, E, device="cuda", dtype=torch.float64)
    refd.index_add_(0, idx.long().T.reshape(-1), dx.double())
    refd[:, F:] = 0
    assert R(dp, refd) < 1e-5
    perms = [[0, 1, 2, 3][:L], {2: [1, 0], 3: [2, 1, 0], 4: [3, 1, 2, 0] if lvl == 3 else [3, 2, 1, 0]}[L]]
    P = ops.make_perms(perms)
    s = ops.perm_concat_fwd(x, P, T, L, E)
    xr = x.view(L, T, E)
    refs = torch.stack([torch.cat([xr[j] for j in pm], dim=-1) for pm in perms]).reshape(2 * T, L * E)
    assert torch.equal(s, refs)
    ds = torch.randn_like(s)
    dxx = ops.perm_concat_bwd(ds, P, T, L, E)
    xg = x.double().requires_grad_(True)
    xr = xg.view(L, T, E)
    torch.stack([torch.cat([xr[j] for j in pm], dim=-1) for pm in perms]).reshape(2 * T, L * E).backward(ds.double())
    assert R(dxx, xg.grad) < 1e-6


def test_featurize_matches_oracle():
    import grappa_oracle as orc
    from grappa_b200 import ops
    g, _ = _random_graph(6)
    names = ["atomic_number", "partial_charge", "ring_encoding", "degree", "charge_model"]
    feats = [g.nodes["n1"].data[k] for k in names]
    out = ops.featurize(feats, g.nodes["n1"].data["partial_charge"], 96)
    ref = orc.input_features(g.cpu(), names)
    assert ref.shape[1] == 85
    assert R(out[:, :85].cpu(), ref) < 1e-6
    assert out[:, 85:].abs().max() == 0


def test_head_outputs_fwd_bwd():
    import grappa_oracle as orc
    from grappa_b200 import ops
    from grappa_b200._lib_ops import HeadOutArgs
    T = 77
    dev = "cuda"
    # bonds
    a = HeadOutArgs(kind=0, T=T, n_perm=2, n_out=2, n_per=0, gated=0, k_mean_over_std=4.7342, k_std=161.2278, k_min=0.0,
                    eq_mean_over_std=6.3251, eq_std=0.1953, eq_min=0.0)
    sc = torch.randn(2 * T, 2, device=dev) * 3
    k, eq = ops.head_output_fwd(a, sc)
    sd = {"p.to_k.std": torch.tensor(161.2278), "p.to_k.mean_over_std": torch.tensor(4.7342), "p.to_k.min_": torch.tensor(0.0),
          "p.to_eq.std": torch.tensor(0.1953), "p.to_eq.mean_over_std": torch.tensor(6.3251), "p.to_eq.min_": torch.tensor(0.0)}
    c = (sc[:T] + sc[T:]).double().cpu().requires_grad_(True)
    rk = orc.to_positive(sd, "p.to_k", c[:, 1])
    req = orc.to_positive(sd, "p.to_eq", c[:, 0])
    assert R(k.cpu(), rk) < 1e-5 and R(eq.cpu(), req) < 1e-5
    dk, deq = torch.randn(T, device=dev), torch.randn(T, device=dev)
    (rk * dk.double().cpu()).sum().backward(retain_graph=True)
    (req * deq.double().cpu()).sum().backward()
    ds = ops.head_output_bwd(a, sc, dk, deq)
    assert R(ds[:T].cpu(), c.grad) < 1e-5 and torch.equal(ds[:T], ds[T:])
    # angles
    a = HeadOutArgs(kind=1, T=T, n_perm=2, n_out=2, k_mean_over_std=3.9726, k_std=26.5965, k_min=0.0,
                    eq_std_over_max=0.029189, eq_max=math.pi)
    k, eq = ops.head_output_fwd(a, sc)
    cc = (sc[:T] + sc[T:]).double().cpu().requires_grad_(True)
    req = math.pi * torch.sigmoid(0.029189 * cc[:, 0])
    assert R(eq.cpu(), req) < 1e-5
    (req * deq.double().cpu()).sum().backward()
    ds = ops.head_output_bwd(a, sc, None, deq)
    assert R(ds[:T, 0].cpu(), cc.grad[:, 0]) < 1e-5 and ds[:, 1].abs().max() == 0
    # gated torsions with cutoff
    a = HeadOutArgs(kind=2, T=T, n_perm=2, n_out=6, n_per=3, gated=1, cutoff=1e-4)
    for i, v in enumerate([0.5977, 1.3465, 0.2466]):
        a.tk_std[i] = v
    sc = torch.randn(2 * T, 6, device=dev)
    sc[5] = 0.0; sc[T + 5] = 0.0           # a tuple that lands below the cutoff
    k, _ = ops.head_output_fwd(a, sc)
    c = (sc[:T] + sc[T:]).double().cpu().requires_grad_(True)
    rk = c[:, :3] * torch.sigmoid(c[:, 3:]) * torch.tensor([0.5977, 1.3465, 0.2466], dtype=torch.float64)
    rk = torch.where(rk.abs() > 1e-4, rk, torch.zeros_like(rk))
    assert R(k.cpu(), rk) < 1e-5 and k[5].abs().max() == 0
    dk = torch.randn(T, 3, device=dev)
    (rk * dk.double().cpu()).sum().backward()
    ds = ops.head_output_bwd(a, sc, dk, None)
    assert R(ds[:T].cpu(), c.grad) < 1e-5


def test_adam_and_clip_match_torch():
    from grappa_b200 import ops
    dev = "cuda"
    n = 100_003
    p = torch.randn(n, device=dev)
    g = torch.randn(n, device=dev) * 3
    ref_p = torch.nn.Parameter(p.clone())
    opt = torch.optim.Adam([ref_p], lr=1e-3)
    m, v = torch.zeros(n, device=dev), torch.zeros(n, device=dev)
    mine = p.clone()
    for step in range(1, 4):
        ref_p.grad = g.clone() * step
        torch.nn.utils.clip_grad_norm_([ref_p], 10.0)
        opt.step()
        gi = g * step
        nsq = torch.zeros(1, device=dev)
        ops.sumsq(gi, nsq)
        ops.adam_step(mine, gi, m, v, 1e-3, 0.9, 0.999, 1e-8, step, gnorm_sq=nsq, clip=10.0)
    assert R(mine, ref_p.data) < 1e-5


def test_molwise_loss_matches_oracle():
    import ctypes as C
    import grappa_oracle as orc
    from grappa_b200 import _lib, synthetic
    from grappa_b200._lib_ops import LossArgs
    from grappa_b200.pack import PackedBatch
    g = synthetic.espaloma_mix_batch(seed=8, batch_size=6, n_confs=5)
    N, Cn, B = g.num_nodes("n1"), 5, 6
    e = torch.randn(B, Cn) * 5
    gr = torch.randn(N, Cn, 3) * 4
    kp = torch.randn(g.num_nodes("n4"), 3)
    ki = torch.randn(g.num_nodes("n4_improper"), 3)
    e64, gr64 = e.double().requires_grad_(True), gr.double().requires_grad_(True)
    kp64, ki64 = kp.double().requires_grad_(True), ki.double().requires_grad_(True)
    ref = orc.molwise_loss({"energy": e64, "gradient": gr64}, {"n4": {"k": kp64}, "n4_improper": {"k": ki64}}, g)
    ref.backward()
    gd = g.to("cuda")
    pack = PackedBatch(gd, device="cuda")
    dev = "cuda"
    t = {k: v.to(dev) for k, v in dict(e=e, gr=gr, kp=kp, ki=ki).items()}
    outs = {k: torch.empty_like(v) for k, v in t.items()}
    loss, mol = torch.zeros(1, device=dev), torch.zeros(B, device=dev)
    a = LossArgs(energy=t["e"].data_ptr(), energy_ref=gd.nodes["g"].data["energy_ref"].data_ptr(), grad=t["gr"].data_ptr(),
                 grad_ref=gd.nodes["n1"].data["gradient_ref"].data_ptr(), atom_off=pack.ptr("atom_off"),
                 k_proper=t["kp"].data_ptr(), k_improper=t["ki"].data_ptr() if ki.numel() else 0,
                 proper_off=pack.ptr("tup_off2"), improper_off=pack.ptr("tup_off3"), B=B, C=Cn, n_per_p=3, n_per_i=3,
                 w_energy=1.0, w_grad=0.8, w_proper=1e-3, w_improper=2e-3, loss=loss.data_ptr(), mol_loss=mol.data_ptr(),
                 g_energy=outs["e"].data_ptr(), g_grad=outs["gr"].data_ptr(), g_k_proper=outs["kp"].data_ptr(),
                 g_k_improper=outs["ki"].data_ptr() if ki.numel() else 0)
    _lib.check(_lib.lib().grappa_b200_molwise_loss(C.byref(a), torch.cuda.current_stream().cuda_stream), "loss")
    assert abs(loss.item() - ref.item()) < 1e-5 * abs(ref.item())
    assert R(outs["e"], e64.grad) < 1e-5 and R(outs["gr"], gr64.grad) < 1e-5 and R(outs["kp"], kp64.grad) < 1e-5
    if ki.numel():
        assert R(outs["ki"], ki64.grad) < 1e-5


# ---------------------------------------------------------------------------------------------------
# tcgen05 / TMA tensor-core GEMM (TF32 operands, fp32 accumulation in TMEM)
# ---------------------------------------------------------------------------------------------------
TC_SHAPES = [(128, 128, 32), (128, 64, 64), (256, 256, 256), (1664, 512, 512), (300, 511, 256), (77, 96, 40),
             (1000, 2048, 512), (130, 1536, 512), (3264, 512, 2048), (64, 16, 8)]


@pytest.mark.parametrize("ta,tb", [(False, False), (False, True), (True, False), (True, True)])
@pytest.mark.parametrize("M,N,K", TC_SHAPES)
def test_gemm_tcgen05_exact_on_tf32_representable_inputs(ta, tb, M, N, K):
    """Small-integer operands are exactly representable in TF32 and the fp32 accumulation is exact,
    so the tensor-core result must equal the fp64 product bit for bit: any descriptor / swizzle /
    layout mistake shows up as a wrong number, not as a tolerance question."""
    from grappa_b200 import ops
    dev = "cuda"
    ld_a = ((M if ta else K) + 3) // 4 * 4
    ld_b = ((N if tb else K) + 3) // 4 * 4
    A = torch.randint(-3, 4, ((K if ta else M), ld_a), device=dev).float()[:, :(M if ta else K)]
    B = torch.randint(-3, 4, ((K if tb else N), ld_b), device=dev).float()[:, :(N if tb else K)]
    ref = ((A.double().T if ta else A.double()) @ (B.double() if tb else B.double().T)).float()
    out = ops.gemm(A, B, trans_a=ta, trans_b=tb, precision=ops.TF32)
    assert torch.equal(out, ref), f"max abs diff {(out - ref).abs().max().item()}"


@pytest.mark.parametrize("ta,tb", [(False, False), (False, True), (True, True)])
def test_gemm_tcgen05_tf32_accuracy_and_epilogue(ta, tb):
    from grappa_b200 import ops
    dev = "cuda"
    M, N, K = 1111, 511, 512
    A = torch.randn((K, 1112) if ta else (M, K), device=dev)[:, :(M if ta else K)]
    B = (torch.randn((K, 512) if tb else (N, K), device=dev) / math.sqrt(K))[:, :(N if tb else K)]
    bias = torch.randn(N, device=dev)
    res = torch.randn(M, N, device=dev)
    ref = F.elu((A.double().T if ta else A.double()) @ (B.double() if tb else B.double().T) + bias.double()) + res.double()
    out = ops.gemm(A, B, trans_a=ta, trans_b=tb, bias=bias, act=1, residual=res, precision=ops.TF32)
    assert R(out, ref) < 1e-3
    same = ops.gemm(A, B, trans_a=ta, trans_b=tb, bias=bias, act=1, residual=res, precision=ops.TF32)
    assert torch.equal(out, same)       # deterministic


def test_gemm_tcgen05_splitk_wgrad():
    from grappa_b200 import ops
    dev = "cuda"
    rows, n_out, k_in = 14848, 512, 512
    dY = torch.randint(-2, 3, (rows, n_out), device=dev).float()
    X = torch.randint(-2, 3, (rows, k_in), device=dev).float()
    ref = (dY.double().T @ X.double()).float()
    out = ops.gemm(dY, X, trans_a=True, trans_b=True, precision=ops.TF32)     # ops.gemm hands in a split-K workspace
    assert torch.equal(out, ref)


def test_gemm_tcgen05_rejects_unaligned_and_auto_falls_back():
    from grappa_b200 import GrappaB200Error, ops
    dev = "cuda"
    A = torch.randn(64, 85, device=dev)          # row pitch 85 floats: not 16-byte aligned -> TMA illegal
    B = torch.randn(32, 85, device=dev)
    with pytest.raises(GrappaB200Error):
        ops.gemm(A, B, precision=ops.TF32)
    out = ops.gemm(A, B, precision=ops.AUTO)     # same call, automatic kernel choice: FFMA path
    assert R(out, A.double() @ B.double().T) < 1e-5


def test_param_loss_term_matches_reference_fixture():
    """Classical-parameter MSE of MolwiseLoss (reference training/loss.py:70-113): value and gradients against the
    fixture the reference's own MolwiseLoss produced (NaN masks, torsion width correction, per-dataset weights)."""
    from grappa_b200.loss import MolwiseLoss
    from util import LEVELS, graph_from_fixture, load_golden
    z = load_golden("param_loss.npz")
    dsnames = [str(d) for d in z["meta.dsnames"]]
    for v in ("a", "b"):
        g = graph_from_fixture(z).to("cuda")
        leaves = {}
        for l in LEVELS:
            for n in ("k", "eq"):
                if f"{v}.in.{l}.{n}" in z.files:
                    t = torch.from_numpy(z[f"{v}.in.{l}.{n}"]).cuda().requires_grad_(True)
                    g.nodes[l].data[n] = t
                    g.nodes[l].data[n + "_ref"] = torch.from_numpy(z[f"{v}.ref.{l}.{n}"]).cuda()
                    leaves[f"{l}.{n}"] = t
        loss = MolwiseLoss(gradient_weight=0., energy_weight=0., param_weight=1e-3,
                           param_weights_by_dataset={"spice": 0.5, "rna": 2.0})(g, dsnames=dsnames)
        assert abs(loss.item() - float(z[f"{v}.loss"])) < 1e-5 * abs(float(z[f"{v}.loss"]))
        keys = [k[len(v) + 6:] for k in z.files if k.startswith(f"{v}.grad.")]
        grads = torch.autograd.grad(loss, [leaves[k] for k in keys])
        for k, gr in zip(keys, grads):
            assert R(gr, torch.from_numpy(z[f"{v}.grad.{k}"])) < 1e-5, k
        uni = MolwiseLoss(gradient_weight=0., energy_weight=0., param_weight=1e-3)(g)
        assert abs(uni.item() - float(z[f"{v}.loss_uniform"])) < 1e-5 * abs(float(z[f"{v}.loss_uniform"]))


@pytest.mark.parametrize("ta,tb", [(False, False), (False, True), (True, False), (True, True)])
@pytest.mark.parametrize("M,N,K", [(1664, 512, 2048), (700, 256, 1024), (300, 1536, 1100)])
def test_gemm_tf32_cta_pair_all_layouts(ta, tb, M, N, K):
    """K >= 1024 selects the cta_group::2 kernel (a cluster of two CTAs per 256-row tile, ragged last pair)."""
    from grappa_b200 import ops
    dev = "cuda"
    A = torch.randn((K, M) if ta else (M, K), device=dev)
    B = torch.randn((K, N) if tb else (N, K), device=dev)
    bias = torch.randn(N, device=dev)
    res = torch.randn(M, N, device=dev)
    ref = (A.double().T if ta else A.double()) @ (B.double() if tb else B.double().T) + bias.double() + res.double()
    out = ops.gemm(A, B, trans_a=ta, trans_b=tb, bias=bias, residual=res, precision=ops.TF32)
    assert R(out, ref) < 1e-3


def test_gemm_grouped_matches_separate_launches():
    """grappa_b200_gemm_grouped == the same problems through grappa_b200_gemm (weight gradients of one layer: different
    output shapes and reduction lengths, split-K partials in one shared workspace, accumulate into an existing gradient)."""
    from grappa_b200 import ops
    dev = "cuda"
    shapes = [(512, 512, 14848), (1536, 512, 14848), (512, 512, 3264), (256, 2048, 1920), (512, 256, 1664)]
    probs, refs = [], []
    for i, (M, N, K) in enumerate(shapes):
        dY = torch.randn(K, M, device=dev)
        X = torch.randn(K, N, device=dev)
        out = torch.randn(M, N, device=dev)
        acc = i % 2 == 1
        ref = ops.gemm(dY, X, trans_a=True, trans_b=True, out=out.clone(), accumulate=acc, precision=ops.TF32)
        exact = dY.double().T @ X.double() + (out.double() if acc else 0)
        assert R(ref, exact) < 1e-3
        probs.append((dY, X, dict(trans_a=True, trans_b=True, out=out, accumulate=acc, precision=ops.TF32)))
        refs.append(exact)
    ops.gemm_grouped(probs)
    for (_, _, kw), exact in zip(probs, refs):
        assert R(kw["out"], exact) < 1e-3
    # fp32 requests fall back to one launch per problem inside the same entry point
    probs32 = [(d, x, dict(kw, out=torch.empty_like(kw["out"]), accumulate=False, precision=ops.FP32)) for d, x, kw in probs[2:]]
    ops.gemm_grouped(probs32)
    for (d, x, kw) in probs32:
        assert R(kw["out"], d.double().T @ x.double()) < 1e-5


@pytest.mark.parametrize("M,N,K,ta,tb", [(7424, 6, 256, False, False), (3264, 2, 256, False, False),
                                         (7424, 256, 6, False, True), (5760, 256, 2, False, True),
                                         (6, 256, 7424, True, True), (2, 256, 3264, True, True)])
def test_skinny_gemm_kernels(M, N, K, ta, tb):
    """Last symmetriser layer (2 / 6 outputs): forward N <= 8, input gradient K <= 8, weight gradient M <= 8."""
    from grappa_b200 import ops
    dev = "cuda"
    A = torch.randn((K, M) if ta else (M, K), device=dev)
    B = torch.randn((K, N) if tb else (N, K), device=dev)
    bias = torch.randn(N, device=dev)
    ref = (A.double().T if ta else A.double()) @ (B.double() if tb else B.double().T) + bias.double()
    for prec in (ops.FP32, ops.AUTO):   # AUTO = what set_matmul_precision('tf32') selects
        out = ops.gemm(A, B, trans_a=ta, trans_b=tb, bias=bias, precision=prec)
        assert R(out, ref) < 1e-5      # these shapes always run on the fp32 streaming kernels


def test_pad_rows_and_padded_pre_dense():
    from grappa_b200 import ops
    dev = "cuda"
    W = torch.randn(512, 85, device=dev)
    Wp = ops.pad_rows(W, 88)
    assert Wp.shape == (512, 88) and torch.equal(Wp[:, :85], W) and Wp[:, 85:].abs().max() == 0
    X = torch.zeros(1664, 88, device=dev)
    X[:, :85] = torch.randn(1664, 85, device=dev)
    ref = X[:, :85].double() @ W.double().T
    got = ops.gemm(X, Wp, k=88, precision=ops.AUTO)
    assert R(got, ref) < 1e-3


def test_loss_ignores_padding_conformations_like_the_reference():
    """Molecules padded to the batch's conformation count ('is_dummy'): value and gradients of the fused loss against
    the fixture produced by the reference's set_number_confs + batch + MolwiseLoss (tests/golden/ragged_confs.npz)."""
    from grappa_b200 import dataset
    from grappa_b200.loss import MolwiseLoss
    from util import graph_from_fixture, load_golden
    z = load_golden("ragged_confs.npz")
    mols = [graph_from_fixture(z, prefix=f"mol{i}.") for i in range(int(z["meta.n_mols"]))]
    for drop_n_valid in (False, True):          # with the collate's n_valid field, and from is_dummy alone
        g = dataset.collate(mols, conf_strategy="max").to("cuda")
        if drop_n_valid:
            del g.nodes["g"].data["n_valid"]
        leaves = {n: torch.from_numpy(z[f"in.{n}"]).cuda().requires_grad_(True) for n in ("energy", "gradient", "k_proper", "k_improper")}
        g.nodes["g"].data["energy"], g.nodes["n1"].data["gradient"] = leaves["energy"], leaves["gradient"]
        g.nodes["n4"].data["k"], g.nodes["n4_improper"].data["k"] = leaves["k_proper"], leaves["k_improper"]
        loss = MolwiseLoss(gradient_weight=0.8, energy_weight=1.0, param_weight=0., proper_regularisation=1e-3,
                           improper_regularisation=1e-3)(g)
        assert abs(loss.item() - float(z["out.loss"])) < 1e-5 * abs(float(z["out.loss"]))
        grads = torch.autograd.grad(loss, list(leaves.values()))
        for n, gr in zip(leaves, grads):
            assert R(gr, torch.from_numpy(z[f"grad.{n}"])) < 1e-5, n
        # padding conformations receive exactly zero gradient
        nv = z["meta.n_confs"].tolist()
        assert all(float(grads[0][b, nv[b]:].abs().sum()) == 0.0 for b in range(len(nv)))


@pytest.mark.parametrize("M,N,K", [(3264, 512, 512), (14848, 512, 512), (700, 256, 256), (1664, 512, 1100)])
def test_gemm_fused_column_sums(M, N, K):
    """dgrad GEMM whose epilogue also emits per-32-row column sums of the stored result (bias-gradient partials)."""
    from grappa_b200 import ops
    dev = "cuda"
    dY = torch.randn(M, K, device=dev)
    W = torch.randn(K, N, device=dev)
    Y = F.elu(torch.randn(M, N, device=dev))
    res = torch.randn(M, N, device=dev)
    out, partial = ops.gemm_with_colsum(dY, W, trans_b=True, residual=res, mul_elu_out=Y, precision=ops.AUTO)
    ref = ops.gemm(dY, W, trans_b=True, residual=res, mul_elu_out=Y, precision=ops.AUTO)
    assert partial is not None and partial.shape == ((M + 31) // 32, N)
    assert torch.equal(out, ref)
    assert R(partial.sum(0), ref.double().sum(0)) < 1e-5
    # row groups are exact sums of their 32 rows
    g = min(5, partial.shape[0] - 1)
    assert R(partial[g], ref[32 * g:32 * g + 32].double().sum(0)) < 1e-5
    # fp32 requests cannot fuse: the helper falls back to a plain GEMM
    out32, p32 = ops.gemm_with_colsum(dY, W, trans_b=True, precision=ops.FP32)
    assert p32 is None and R(out32, dY.double() @ W.double()) < 1e-5


# ---------------------------------------------------------------------------------------------------
# bf16x3 tensor-core GEMM: fp32 operands split in-kernel into bf16 hi + lo, hi*hi + lo*hi + hi*lo, fp32 accumulation
# ---------------------------------------------------------------------------------------------------
X3_TOL = 2e-5      # 16 operand mantissa bits + the dropped lo*lo term; TF32 sits at ~5e-4 on the same inputs


@pytest.mark.parametrize("ta,tb", [(False, False), (False, True), (True, False), (True, True)])
@pytest.mark.parametrize("M,N,K", TC_SHAPES)
def test_gemm_bf16x3_exact_on_small_integers(ta, tb, M, N, K):
    """Small integers are exact in bf16 (lo part = 0) and fp32 accumulation is exact: bit-equality with the fp64 product
    pins the converter's layout (swizzle, transposition of MN-major operands) and the MMA descriptors."""
    from grappa_b200 import ops
    dev = "cuda"
    ld_a = ((M if ta else K) + 3) // 4 * 4
    ld_b = ((N if tb else K) + 3) // 4 * 4
    A = torch.randint(-3, 4, ((K if ta else M), ld_a), device=dev).float()[:, :(M if ta else K)]
    B = torch.randint(-3, 4, ((K if tb else N), ld_b), device=dev).float()[:, :(N if tb else K)]
    ref = ((A.double().T if ta else A.double()) @ (B.double() if tb else B.double().T)).float()
    out = ops.gemm(A, B, trans_a=ta, trans_b=tb, precision=ops.BF16X3)
    assert torch.equal(out, ref), f"max abs diff {(out - ref).abs().max().item()}"


@pytest.mark.parametrize("ta,tb", [(False, False), (False, True), (True, False), (True, True)])
def test_gemm_bf16x3_hi_lo_terms_exact(ta, tb):
    """a = ah + al with ah in {-2..2} and al in {-3..3} * 2^-10: hi = ah, lo = al exactly; with B integer-valued the
    product needs exactly hi*hi + lo*hi (and with the roles swapped hi*lo): every cross term is checked bit for bit."""
    from grappa_b200 import ops
    dev = "cuda"
    M, N, K = 300, 192, 160
    def two_part(shape):
        ah = torch.randint(-2, 3, shape, device=dev).float()
        ah = torch.where(ah == 0, torch.ones_like(ah), ah)                      # hi part never zero, so lo stays the lo part
        return ah + torch.randint(-3, 4, shape, device=dev).float() * 2.0 ** -10
    ints = lambda shape: torch.randint(-3, 4, shape, device=dev).float()
    for split_a in (True, False):
        A = (two_part if split_a else ints)((K, M) if ta else (M, K))
        B = (ints if split_a else two_part)((K, N) if tb else (N, K))
        ref = ((A.double().T if ta else A.double()) @ (B.double() if tb else B.double().T)).float()
        out = ops.gemm(A, B, trans_a=ta, trans_b=tb, precision=ops.BF16X3)
        assert torch.equal(out, ref), (split_a, (out - ref).abs().max().item())


@pytest.mark.parametrize("ta,tb", [(False, False), (False, True), (True, False), (True, True)])
@pytest.mark.parametrize("M,N,K", [(1111, 511, 512), (1664, 512, 2048), (700, 256, 1024), (300, 1536, 1100), (14848, 512, 512), (64, 16, 8)])
def test_gemm_bf16x3_accuracy_epilogue_and_pairs(ta, tb, M, N, K):
    """Random operands against fp64 at fp32-class tolerance (single-CTA tiles, CTA pairs for K >= 1024, ragged edges,
    fused bias / ELU / residual epilogue), deterministic."""
    from grappa_b200 import ops
    dev = "cuda"
    pa, pb = ((M if ta else K) + 3) // 4 * 4, ((N if tb else K) + 3) // 4 * 4
    A = torch.randn((K if ta else M), pa, device=dev)[:, :(M if ta else K)]
    B = (torch.randn((K if tb else N), pb, device=dev) / math.sqrt(K))[:, :(N if tb else K)]
    bias = torch.randn(N, device=dev)
    res = torch.randn(M, N, device=dev)
    pre = (A.double().T if ta else A.double()) @ (B.double() if tb else B.double().T)
    out = ops.gemm(A, B, trans_a=ta, trans_b=tb, bias=bias, act=1, residual=res, precision=ops.BF16X3)
    assert R(out, F.elu(pre + bias.double()) + res.double()) < X3_TOL
    plain = ops.gemm(A, B, trans_a=ta, trans_b=tb, precision=ops.BF16X3)
    err = R(plain, pre)
    tf = R(ops.gemm(A, B, trans_a=ta, trans_b=tb, precision=ops.AUTO), pre)
    print(f"bf16x3 {err:.2e} vs tf32 {tf:.2e}")
    assert err < X3_TOL
    assert torch.equal(plain, ops.gemm(A, B, trans_a=ta, trans_b=tb, precision=ops.BF16X3))


def test_gemm_bf16x3_splitk_grouped_and_column_sums():
    from grappa_b200 import ops
    dev = "cuda"
    # split-K weight gradient (ops.gemm hands in the workspace), exact on integers
    dY = torch.randint(-2, 3, (14848, 512), device=dev).float()
    X = torch.randint(-2, 3, (14848, 512), device=dev).float()
    assert torch.equal(ops.gemm(dY, X, trans_a=True, trans_b=True, precision=ops.BF16X3), (dY.double().T @ X.double()).float())
    # grouped launch of four weight gradients
    shapes = [(512, 512, 14848), (1536, 512, 14848), (512, 512, 3264), (256, 2048, 1920), (512, 256, 1664)]
    probs, refs = [], []
    for i, (M, N, K) in enumerate(shapes):
        dY = torch.randn(K, M, device=dev)
        X = torch.randn(K, N, device=dev)
        out = torch.randn(M, N, device=dev)
        acc = i % 2 == 1
        refs.append(dY.double().T @ X.double() + (out.double() if acc else 0))
        probs.append((dY, X, dict(trans_a=True, trans_b=True, out=out, accumulate=acc, precision=ops.BF16X3)))
    ops.gemm_grouped(probs)
    for (_, _, kw), exact in zip(probs, refs):
        assert R(kw["out"], exact) < X3_TOL
    # fused column sums in the dgrad epilogue
    M, N, K = 3264, 512, 512
    dY = torch.randn(M, K, device=dev)
    W = torch.randn(K, N, device=dev)
    Y = F.elu(torch.randn(M, N, device=dev))
    res = torch.randn(M, N, device=dev)
    out, partial = ops.gemm_with_colsum(dY, W, trans_b=True, residual=res, mul_elu_out=Y, precision=ops.BF16X3)
    exact = (dY.double() @ W.double()) * torch.where(Y > 0, torch.ones_like(Y), Y + 1).double() + res.double()
    assert partial is not None and R(out, exact) < X3_TOL and R(partial.sum(0), out.double().sum(0)) < 1e-5
    # not TMA-legal -> FFMA fallback, still correct
    A = torch.randn(64, 85, device=dev)
    B = torch.randn(32, 85, device=dev)
    assert R(ops.gemm(A, B, precision=ops.BF16X3), A.double() @ B.double().T) < 1e-5
