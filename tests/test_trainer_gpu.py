"""Training step on the GPU: the captured-CUDA-graph path must reproduce the eager path, keep drawing fresh dropout
masks on replay, and follow torch.optim.Adam + clip_grad_norm_ (reference training/lightning_model.py:297-299,
config.py:106)."""
import copy

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _setup(dropout: float, seed: int = 5, **switches):
    import grappa_oracle as orc
    from grappa_b200 import models, ops, synthetic
    from grappa_b200.energy import Energy
    from grappa_b200.loss import MolwiseLoss
    ops.set_matmul_precision("fp32")
    cfg = dict(orc.small_model_config())
    for k in ("gnn_dropout_attention", "gnn_dropout_initial", "gnn_dropout_final", "parameter_dropout"):
        cfg[k] = dropout
    cfg.update(switches)
    model = models.model_from_config(cfg)
    model.load_state_dict(synthetic.deterministic_state_dict(model.state_dict(), seed=seed))
    loss = MolwiseLoss(gradient_weight=0.8, energy_weight=1.0, param_weight=0.0, proper_regularisation=1e-3,
                       improper_regularisation=1e-3)
    return model.train(), Energy(write_tuple_terms=False), loss


def _batches(n):
    from grappa_b200 import synthetic
    from grappa_b200.pack import get_pack
    out = []
    for i in range(n):
        g = synthetic.peptide_batch(seed=20 + i, batch_size=3, n_res=1, n_confs=5)
        get_pack(g)
        out.append(g)
    return out


def test_graph_replay_equals_eager_steps():
    from grappa_b200.training import Trainer
    batches = _batches(5)
    results = []
    for use_graph in (False, True):
        model, energy, loss = _setup(0.0)
        tr = Trainer(model, energy, loss, lr=1e-3, clip=10.0, device="cuda", use_cuda_graph=use_graph)
        losses = [float(tr.step(g).item()) for g in batches]
        results.append((losses, tr.fp.flat.detach().cpu().numpy().copy(), tr))
    (l0, p0, _), (l1, p1, tr) = results
    assert len(tr._captured) == 1, "same-shaped batches must share one captured graph"
    assert l1 == l0, (l0, l1)
    # every kernel of the step is deterministic (fixed-order split-K / column / norm reductions, round-scheduled force
    # accumulation), so the captured multi-stream graph must reproduce the eager single-launch path BIT FOR BIT --
    # any difference would mean a missing dependency between the streams of the graph
    d = np.abs(p1 - p0)
    print("graph vs eager: max |dp| =", d.max())
    assert d.max() == 0.0, d.max()
    assert l0[-1] != l0[0]


def test_learnable_statistics_train_inside_the_captured_step():
    """learnable_statistics=True: the statistics live in the flat parameter buffer, the head-output kernels read them
    from there, so graph replays must see each Adam update -- same losses and bit-identical parameters as eager steps,
    and the statistics really move."""
    from grappa_b200.training import Trainer
    batches = _batches(4)
    results = []
    for use_graph in (False, True):
        model, energy, loss = _setup(0.0, learnable_statistics=True, layer_norm=False)
        tr = Trainer(model, energy, loss, lr=1e-3, clip=10.0, device="cuda", use_cuda_graph=use_graph)
        std0 = float(model.parameter_writer.bond_writer.to_eq.std)
        losses = [float(tr.step(g).item()) for g in batches]
        stats = [float(model.parameter_writer.bond_writer.to_eq.std), float(model.parameter_writer.angle_writer.to_eq.std_over_max),
                 model.parameter_writer.proper_writer.k_std.detach().cpu().numpy().copy()]
        assert abs(stats[0] - std0) > 1e-4, "the statistic did not train"
        results.append((losses, tr.fp.flat.detach().cpu().numpy().copy(), stats))
    (l0, p0, s0), (l1, p1, s1) = results
    assert l1 == l0, (l0, l1)
    assert np.abs(p1 - p0).max() == 0.0
    assert s0[0] == s1[0] and s0[1] == s1[1] and np.array_equal(s0[2], s1[2])


def test_graph_replay_draws_fresh_dropout_masks_and_honours_lr():
    from grappa_b200.training import Trainer
    g = _batches(1)[0]
    model, energy, loss = _setup(0.3)
    tr = Trainer(model, energy, loss, lr=0.0, clip=10.0, device="cuda", use_cuda_graph=True)
    losses = [float(tr.step(g).item()) for _ in range(5)]     # lr = 0: parameters frozen, only the masks change
    assert len(tr._captured) == 1
    assert len(set(losses[1:])) == len(losses[1:]), f"replays must not repeat the dropout mask: {losses}"
    before = tr.fp.flat.clone()
    tr.lr = 1e-3
    tr.step(g)
    assert (tr.fp.flat - before).abs().max().item() > 0
    assert int(tr.counters[0].item()) == 6 and tr.step_count == 6


def test_adam_and_clip_match_torch():
    """One eager step against torch.optim.Adam + clip_grad_norm_ on the same gradients."""
    from grappa_b200.training import Trainer
    g = _batches(1)[0].to("cuda")
    model, energy, loss = _setup(0.0)
    ref = copy.deepcopy(model).cuda()
    tr = Trainer(model, energy, loss, lr=1e-3, clip=0.05, device="cuda", use_cuda_graph=False)
    p_before = tr.fp.flat.clone()
    tr.forward_backward(g)
    grads = tr.fp.grad.clone()
    tr.optimizer_step()
    # torch reference on the identical gradient vector
    p = torch.nn.Parameter(p_before.clone())
    p.grad = grads.clone()
    opt = torch.optim.Adam([p], lr=1e-3)
    total = torch.nn.utils.clip_grad_norm_([p], 0.05)
    assert total.item() > 0.05, "the clip must be active for this test to mean something"
    opt.step()
    np.testing.assert_allclose(tr.fp.flat.cpu().numpy(), p.detach().cpu().numpy(), rtol=2e-6, atol=1e-9)


def test_training_from_packed_dataset_through_the_prefetch_loader():
    """dataset.PackedDataset -> PrefetchLoader (worker thread, pinned batches, ragged conformation counts) ->
    Trainer.step on host graphs: losses stay finite, batches with a new shape run eagerly first and are captured later."""
    import numpy as np
    from grappa_b200 import dataset, synthetic
    from grappa_b200.training import Trainer
    model, energy, loss = _setup(0.1)
    tr = Trainer(model, energy, loss, lr=1e-3, clip=10.0, device="cuda", use_cuda_graph=True)
    rng = np.random.default_rng(0)
    mols = []
    for i in range(12):
        m = synthetic.make_molecule(rng, "peptide", n_confs=int(rng.integers(2, 7)), n_res=1 + i % 2)
        if m.num_nodes("n4_improper") > 0:
            mols.append(m)
    ds = dataset.PackedDataset.from_graphs(mols, dsnames=["a", "b"] * (len(mols) // 2) + ["a"] * (len(mols) % 2))
    losses = []
    for epoch in range(3):
        batches = list(dataset.batch_sampler(range(len(ds)), 4, np.random.default_rng(epoch)))
        for g in dataset.PrefetchLoader(ds, batches, conf_strategy=4, seed=epoch, depth=2):
            assert g.nodes["g"].data["is_dummy"].shape == (4, g.nodes["n1"].data["xyz"].shape[1])
            losses.append(tr.step(g).item())
    assert len(losses) == 3 * (len(ds) // 4) and all(np.isfinite(losses))


def test_captured_step_keeps_padding_and_param_weights():
    """The per-batch loss inputs `n_valid` / `is_dummy` (padded conformations) and `param_weight` (per-dataset weights of the
    classical-parameter term) are inputs of the captured step like xyz: a replay must optimise exactly the loss the eager
    step optimises -- same losses, bit-identical parameters.  (Round 1 dropped them from the static graph: replays counted
    the padded conformations and fell back to the uniform weight.)"""
    from grappa_b200 import dataset, synthetic
    from grappa_b200.loss import MolwiseLoss
    from grappa_b200.training import Trainer
    rng = np.random.default_rng(3)
    mols = []
    for i in range(12):
        m = synthetic.make_molecule(rng, "peptide", n_confs=(2, 3, 4, 6)[i % 4], n_res=1)     # one topology, ragged confs
        for lvl, names in (("n2", ("k", "eq")), ("n3", ("k", "eq")), ("n4", ("k",))):
            T = m.num_nodes(lvl)
            for n in names:
                shape = (T, 3) if lvl == "n4" else (T,)
                m.nodes[lvl].data[n + "_ref"] = torch.from_numpy(rng.normal(1.0, 0.3, size=shape).astype(np.float32))
        mols.append(m)
    ds = dataset.PackedDataset.from_graphs(mols, dsnames=["a", "b", "c"] * 4)
    batches = [[0, 1, 2, 3], [4, 5, 6, 7], [8, 9, 10, 11], [3, 7, 11, 0], [1, 5, 9, 2]]

    def run(use_graph):
        model, energy, _ = _setup(0.0)
        loss = MolwiseLoss(gradient_weight=0.8, energy_weight=1.0, param_weight=1e-3, proper_regularisation=1e-3,
                           improper_regularisation=1e-3)
        tr = Trainer(model, energy, loss, lr=1e-3, clip=10.0, device="cuda", use_cuda_graph=use_graph)
        losses = []
        for g in dataset.PrefetchLoader(ds, batches, conf_strategy=5, seed=1, depth=2, param_weight=1e-3,
                                        param_weights_by_dataset={"a": 0.0, "b": 5e-2}):
            assert float(g.nodes["g"].data["is_dummy"].sum()) > 0, "the test needs padded conformations"
            losses.append(float(tr.step(g).item()))
        return losses, tr.fp.flat.detach().cpu().numpy().copy(), tr

    l0, p0, _ = run(False)
    l1, p1, tr = run(True)
    assert len(tr._captured) == 1, "same-shaped batches must share one captured graph"
    keys = {k for _, k in next(iter(tr._captured.values())).keys}
    assert {"n_valid", "is_dummy", "param_weight"} <= keys, keys
    assert l1 == l0, (l0, l1)
    assert np.abs(p1 - p0).max() == 0.0
    # and the fields matter: ignoring them changes the loss of the same batch
    model, energy, _ = _setup(0.0)
    loss = MolwiseLoss(gradient_weight=0.8, energy_weight=1.0, param_weight=1e-3, proper_regularisation=1e-3,
                       improper_regularisation=1e-3)
    g = next(iter(dataset.PrefetchLoader(ds, batches[:1], conf_strategy=5, seed=1, param_weight=1e-3,
                                         param_weights_by_dataset={"a": 0.0, "b": 5e-2}))).to("cuda")
    full = float(loss(energy(model.cuda()(g))).item())
    for k in ("n_valid", "is_dummy", "param_weight"):
        del g.nodes["g"].data[k]
    stripped = float(loss(energy(model(g))).item())
    assert abs(full - stripped) > 1e-6 * abs(full)
