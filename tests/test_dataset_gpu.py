"""Device-side batch assembly (dataset.DeviceDataset, kernel grappa_b200_collate) against the host collate
(dataset.PackedDataset.collate, itself pinned against the reference's set_number_confs + batch in tests/test_dataset.py):
every graph field and every index table of the PackedBatch must be bit-identical, for ragged conformation counts
(sub-sampling and padding), mixed molecule sizes, and through a training step."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _dataset(n=14, seed=0):
    from grappa_b200 import dataset, synthetic
    rng = np.random.default_rng(seed)
    mols = []
    kinds = ("peptide", "small", "peptide", "rna", "small")
    while len(mols) < n:
        kind = kinds[len(mols) % len(kinds)]
        kw = {"n_res": 1 + len(mols) % 3} if kind == "peptide" else ({"n_atoms": int(rng.integers(5, 40))} if kind == "small" else {})
        m = synthetic.make_molecule(rng, kind, n_confs=int(rng.integers(2, 9)), **kw)
        if m.num_nodes("n4_improper") > 0:
            mols.append(m)
    return dataset.PackedDataset.from_graphs(mols, dsnames=["a", "b"] * (n // 2))


@pytest.mark.parametrize("conf_strategy", [4, "max", "min", 100])
def test_device_collate_equals_host_collate(conf_strategy):
    from grappa_b200 import dataset
    from grappa_b200.pack import get_pack
    ds = _dataset()
    dd = dataset.DeviceDataset(ds, "cuda")
    for trial, idx in enumerate(([0, 1, 2, 3], [13, 5, 5, 7, 2, 9], [4], list(range(14)))):
        host = ds.collate(idx, conf_strategy, np.random.default_rng([7, trial]))
        devg = dd.collate(idx, conf_strategy, np.random.default_rng([7, trial]))
        torch.cuda.synchronize()
        assert devg.device.type == "cuda"
        for nt in host.ntypes:
            assert devg.num_nodes(nt) == host.num_nodes(nt)
            assert torch.equal(devg.batch_num_nodes(nt).cpu(), host.batch_num_nodes(nt))
            assert set(devg.nodes[nt].data.keys()) == set(host.nodes[nt].data.keys()), nt
            for k, v in host.nodes[nt].data.items():
                got = devg.nodes[nt].data[k]
                assert got.dtype == v.dtype and tuple(got.shape) == tuple(v.shape), (nt, k, got.dtype, got.shape, v.shape)
                assert torch.equal(got.cpu(), v), (nt, k)
        hs, hd = host.edges()
        gs, gd = devg.edges()
        assert torch.equal(gs.cpu(), hs) and torch.equal(gd.cpu(), hd)
        ph, pd = get_pack(host), get_pack(devg)
        assert pd.signature() == ph.to("cuda").signature()
        for name in ph._names:
            a, b = ph.host[name], pd[name].cpu().numpy()
            assert np.array_equal(a.reshape(-1), b.reshape(-1)), name


def test_training_step_from_device_batches_equals_host_batches():
    """Trainer.step on batches assembled on the device == on the host-collated batches (same sampler / RNG streams):
    identical losses and bit-identical parameters, eager and captured."""
    import grappa_oracle as orc
    from grappa_b200 import dataset, models, ops, synthetic
    from grappa_b200.energy import Energy
    from grappa_b200.loss import MolwiseLoss
    from grappa_b200.training import Trainer
    ds = _dataset(12, seed=3)
    dd = dataset.DeviceDataset(ds, "cuda")
    batches = [[0, 1, 2, 3], [4, 5, 6, 7], [8, 9, 10, 11], [0, 1, 2, 3], [4, 5, 6, 7]]
    results = []
    for source in ("host", "device"):
        ops.set_matmul_precision("fp32")
        cfg = dict(orc.small_model_config())
        for k in ("gnn_dropout_attention", "gnn_dropout_initial", "gnn_dropout_final", "parameter_dropout"):
            cfg[k] = 0.0
        model = models.model_from_config(cfg)
        model.load_state_dict(synthetic.deterministic_state_dict(model.state_dict(), seed=5))
        loss = MolwiseLoss(gradient_weight=0.8, energy_weight=1.0, param_weight=0.0, proper_regularisation=1e-3,
                           improper_regularisation=1e-3)
        tr = Trainer(model.train(), Energy(write_tuple_terms=False), loss, lr=1e-3, clip=10.0, device="cuda", use_cuda_graph=True)
        losses = []
        for k, idx in enumerate(batches):
            rng = np.random.default_rng([11, k])
            g = ds.collate(idx, 4, rng) if source == "host" else dd.collate(idx, 4, rng)
            losses.append(float(tr.step(g).item()))
        results.append((losses, tr.fp.flat.detach().cpu().numpy().copy()))
    (l0, p0), (l1, p1) = results
    assert l0 == l1, (l0, l1)
    assert np.abs(p0 - p1).max() == 0.0
