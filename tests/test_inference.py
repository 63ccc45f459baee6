"""Inference front-end: topology -> graph, Parameters conventions, batched prediction."""
import numpy as np
import pytest
import torch


def test_molecule_graph_from_topology_and_parameter_conventions():
    from grappa_b200 import inference, synthetic
    el, bonds, imp = synthetic.polyalanine_topology(2)
    n = len(el)
    ids = (np.arange(n) * 3 + 100).astype(np.int64)                  # external atom ids != indices
    g = inference.molecule_graph(ids, ids[np.asarray(bonds)], el, np.linspace(-0.5, 0.5, n), ids[np.asarray(imp)].reshape(-1, 4))
    ref = synthetic.make_molecule(np.random.default_rng(0), "peptide", n_confs=0, n_res=2)
    for lvl in ("n2", "n3", "n4", "n4_improper"):
        assert torch.equal(g.nodes[lvl].data["idxs"], ref.nodes[lvl].data["idxs"]), lvl
    for f in ("atomic_number", "ring_encoding", "degree", "charge_model"):
        assert torch.equal(g.nodes["n1"].data[f], ref.nodes["n1"].data[f]), f
    assert inference.connected_components(g) == 1
    # Parameters.from_dgl conventions (reference data/Parameters.py:63-140)
    rng = np.random.default_rng(1)
    g.nodes["n2"].data["k"] = torch.full((g.num_nodes("n2"),), 500.0); g.nodes["n2"].data["eq"] = torch.full((g.num_nodes("n2"),), 1.2)
    g.nodes["n3"].data["k"] = torch.full((g.num_nodes("n3"),), 80.0); g.nodes["n3"].data["eq"] = torch.full((g.num_nodes("n3"),), 1.9)
    kp = torch.from_numpy(rng.normal(size=(g.num_nodes("n4"), 3)).astype(np.float32)); kp[0, 0] = 0.0
    ki = torch.from_numpy(rng.normal(size=(g.num_nodes("n4_improper"), 3)).astype(np.float32)); ki[0, 0] = 0.0
    g.nodes["n4"].data["k"], g.nodes["n4_improper"].data["k"] = kp, ki
    p = inference.Parameters.from_graph(g)
    assert np.array_equal(p.atoms, ids) and np.array_equal(p.bonds, ids[ref.nodes["n2"].data["idxs"].numpy()])
    assert np.all(p.proper_ks >= 0) and np.allclose(p.proper_ks * np.cos(p.proper_phases), kp.numpy(), atol=1e-6)
    assert np.allclose(p.improper_ks * np.cos(p.improper_phases), ki.numpy(), atol=1e-6)
    assert p.proper_phases[0, 0] == 0.0 and p.improper_phases[0, 0] == np.float32(np.pi)   # k == 0: '>=' vs '>' as in the reference
    g.nodes["n2"].data["eq"][3] = 0.2
    with pytest.raises(RuntimeError):
        inference.Parameters.from_graph(g)
    inference.Parameters.from_graph(g, check_eq_values=False)
    # two molecules in one graph are reported as disconnected
    from grappa_b200 import graph as gbg
    assert inference.connected_components(gbg.batch([ref, ref])) == 2
    assert inference.shard(list(range(7)), 1, 3) == [1, 4]


@pytest.mark.gpu
def test_predict_many_matches_single_molecule_calls():
    import grappa_oracle as orc
    from grappa_b200 import inference, models, ops, synthetic
    ops.set_matmul_precision("fp32")
    model = models.model_from_config(dict(orc.small_model_config()))
    model.load_state_dict(synthetic.deterministic_state_dict(model.state_dict(), seed=2))
    gr = inference.Grappa(model, device="cuda")
    rng = np.random.default_rng(5)
    mols = [synthetic.make_molecule(rng, k, n_confs=0, **kw) for k, kw in
            (("peptide", dict(n_res=3)), ("small", dict(n_atoms=11)), ("rna", {}), ("peptide", dict(n_res=1)))]
    many = gr.predict_many(mols, max_atoms_per_batch=120, check_eq_values=False)       # forces several model calls
    assert len(many) == len(mols)
    for m, p in zip(mols, many):
        single = gr.predict(m, check_eq_values=False)
        assert p.bonds.shape == (m.num_nodes("n2"), 2) and p.improper_ks.shape == (m.num_nodes("n4_improper"), 3)
        for f in ("bond_k", "bond_eq", "angle_k", "angle_eq", "proper_ks", "improper_ks", "proper_phases"):
            a, b = getattr(p, f), getattr(single, f)
            assert np.max(np.abs(a - b)) <= 1e-5 * max(1e-30, np.max(np.abs(b))), f
    with pytest.raises(ValueError):
        from grappa_b200 import graph as gbg
        gr.predict_many([gbg.batch([mols[1], mols[3]])])


@pytest.mark.gpu
def test_graph_replay_inference_equals_eager():
    """Second and later batches of a known shape replay a captured forward: identical parameters, fresh inputs honoured."""
    import grappa_oracle as orc
    from grappa_b200 import inference, models, ops, synthetic
    ops.set_matmul_precision("fp32")
    model = models.model_from_config(dict(orc.small_model_config()))
    model.load_state_dict(synthetic.deterministic_state_dict(model.state_dict(), seed=3))
    eager = inference.Grappa(model, device="cuda", use_cuda_graph=False)
    graphed = inference.Grappa(model, device="cuda", use_cuda_graph=True)
    rng = np.random.default_rng(9)
    base = synthetic.make_molecule(rng, "peptide", n_confs=0, n_res=2)
    variants = []
    for i in range(4):                       # same topology (same shape signature), different charges
        m = synthetic.make_molecule(np.random.default_rng(9), "peptide", n_confs=0, n_res=2)
        m.nodes["n1"].data["partial_charge"] = m.nodes["n1"].data["partial_charge"] + 0.05 * i
        variants.append(m)
    for i, m in enumerate(variants):
        a = eager.predict(m, check_eq_values=False)
        b = graphed.predict(m, check_eq_values=False)
        for f in ("bond_k", "bond_eq", "angle_k", "angle_eq", "proper_ks", "improper_ks"):
            assert np.array_equal(getattr(a, f), getattr(b, f)), (i, f)
        assert np.array_equal(a.propers, b.propers)
    assert len(graphed._captured) == 1
    assert not np.array_equal(graphed.predict(variants[0], check_eq_values=False).bond_k,
                              graphed.predict(variants[3], check_eq_values=False).bond_k)


def test_parameters_from_graph_matches_the_reference_class():
    """`Parameters.from_graph` against the REFERENCE's own `Parameters.from_dgl` (data/Parameters.py:63-140), run in the
    build container on the reference model's outputs for a molecule with non-trivial atom ids
    (tests/golden/dropin_tiny.npz, generated by tests/golden/make_golden.py::dropin_case)."""
    from grappa_b200 import graph as gbg, inference
    from util import LEVELS, graph_from_fixture, load_golden
    z = load_golden("dropin_tiny.npz")
    g = graph_from_fixture(z)
    for l in LEVELS:                       # the reference's parameter outputs, written where the model writes them
        g.nodes[l].data["k"] = torch.from_numpy(z[f"out.{l}.k"])
        if l in ("n2", "n3"):
            g.nodes[l].data["eq"] = torch.from_numpy(z[f"out.{l}.eq"])
    m0 = gbg.unbatch(g)[0]
    p = inference.Parameters.from_graph(m0)
    for f in ("atoms", "bonds", "angles", "propers", "impropers"):
        assert np.array_equal(getattr(p, f), z[f"params.{f}"]), f
    for f in ("bond_k", "bond_eq", "angle_k", "angle_eq", "proper_ks", "improper_ks"):
        # the fixture's Parameters come from a single-molecule model call, k / eq above from the batched one: fp32 rounding
        np.testing.assert_allclose(getattr(p, f), z[f"params.{f}"], rtol=2e-5, atol=2e-6, err_msg=f)
    for f, ks in (("proper_phases", "proper_ks"), ("improper_phases", "improper_ks")):
        settled = z[f"params.{ks}"] > 1e-5                     # an amplitude at the cutoff may round to either side of zero
        assert np.array_equal(getattr(p, f)[settled], z[f"params.{f}"][settled]), f
