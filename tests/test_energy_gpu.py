"""K13 / K14 parity: CUDA energy + forces + parameter gradients vs the reference-generated fixture
(tests/golden/energy_mixed_batch.npz) and vs the fp64 oracle, through the C ABI.

Tolerances (BASELINE.json north_star): energies / forces 1e-5 relative in fp32, gradients 1e-4.
"""
import numpy as np
import pytest
import torch

from util import LEVELS, graph_from_fixture, load_golden, rel_err

pytestmark = pytest.mark.gpu

TOL_E = 1e-5
TOL_G = 1e-4


def _graph_with_params(z, dev, requires_grad=False):
    g = graph_from_fixture(z).to(dev)
    leaves = {}
    for l in LEVELS:
        for n in ("k", "eq"):
            key = f"in.{l}.{n}"
            if key in z.files:
                t = torch.from_numpy(z[key]).to(dev).requires_grad_(requires_grad)
                g.nodes[l].data[n] = t
                leaves[(l, n)] = t
    return g, leaves


@pytest.mark.parametrize("variant", [1, 2, 3, 4, 5])
def test_energy_forward_matches_reference_fixture(variant):
    from grappa_b200.energy import Energy
    z = load_golden("energy_mixed_batch.npz")
    g, _ = _graph_with_params(z, "cuda")
    en = Energy()
    en.kernel_variant = variant
    if variant == 2:
        # the fixture batch has a 96-atom rna-like molecule: fits the tiled kernel
        pass
    with torch.no_grad():
        g = en(g)
    torch.cuda.synchronize()
    assert rel_err(g.nodes["g"].data["energy"].cpu().numpy(), z["out.g.energy"]) < TOL_E
    assert rel_err(g.nodes["n1"].data["gradient"].cpu().numpy(), z["out.n1.gradient"]) < TOL_E
    for l in LEVELS:
        assert rel_err(g.nodes["g"].data[f"energy_{l}"].cpu().numpy(), z[f"out.g.energy_{l}"]) < TOL_E
        x = g.nodes[l].data["x"].cpu().numpy()
        ref = z[f"out.{l}.x"]
        if l in ("n4", "n4_improper"):
            d = np.abs(np.angle(np.exp(1j * (x - ref))))     # compare angles modulo 2 pi
            assert d.max() < 2e-5
        else:
            assert rel_err(x, ref) < TOL_E


def test_energy_backward_matches_reference_double_backward():
    from grappa_b200.energy import Energy
    z = load_golden("energy_mixed_batch.npz")
    g, leaves = _graph_with_params(z, "cuda", requires_grad=True)
    g = Energy(write_tuple_terms=False)(g)
    gE = torch.from_numpy(z["in.gE"]).cuda()
    gF = torch.from_numpy(z["in.gF"]).cuda()
    obj = (g.nodes["g"].data["energy"] * gE).sum() + (g.nodes["n1"].data["gradient"] * gF).sum()
    obj.backward()
    for (l, n), t in leaves.items():
        assert rel_err(t.grad.cpu().numpy(), z[f"grad.{l}.{n}"]) < TOL_G, (l, n)


def test_energy_vs_fp64_oracle_and_finite_differences():
    """Independent cross-check: fp64 oracle energies; forces = central finite differences of E."""
    import grappa_oracle as orc
    from grappa_b200 import synthetic
    from grappa_b200.energy import Energy
    g = synthetic.peptide_batch(seed=3, batch_size=4, n_res=2, n_confs=5)
    gen = torch.Generator().manual_seed(0)
    prm = {}
    for l in LEVELS:
        T = g.num_nodes(l)
        if l == "n2":
            prm[l] = {"k": 400 + 200 * torch.rand(T, generator=gen), "eq": 1.0 + 0.4 * torch.rand(T, generator=gen)}
        elif l == "n3":
            prm[l] = {"k": 50 + 50 * torch.rand(T, generator=gen), "eq": 1.8 + 0.3 * torch.rand(T, generator=gen)}
        else:
            prm[l] = {"k": torch.randn(T, 3, generator=gen)}
        for n, v in prm[l].items():
            g.nodes[l].data[n] = v
    idxs = {l: g.nodes[l].data["idxs"] for l in LEVELS}
    counts = {l: g.batch_num_nodes(l).tolist() for l in LEVELS}
    p64 = {l: {n: v.double() for n, v in d.items()} for l, d in prm.items()}
    xyz64 = g.nodes["n1"].data["xyz"].double()
    ref = orc.energy_forward(xyz64, idxs, p64, counts)
    gd = Energy()(g.to("cuda"))
    assert rel_err(gd.nodes["g"].data["energy"].cpu().numpy(), ref["energy"].detach().numpy()) < TOL_E
    assert rel_err(gd.nodes["n1"].data["gradient"].cpu().numpy(), ref["gradient"].numpy()) < TOL_E
    # finite differences on one coordinate of the fp64 oracle agree with its autograd gradient
    h = 1e-5
    xp, xm = xyz64.clone(), xyz64.clone()
    xp[3, 2, 1] += h
    xm[3, 2, 1] -= h
    ep = orc.energy_forward(xp, idxs, p64, counts, gradients=False)["energy"].sum()
    em = orc.energy_forward(xm, idxs, p64, counts, gradients=False)["energy"].sum()
    fd = float((ep - em) / (2 * h))
    assert abs(fd - float(gd.nodes["n1"].data["gradient"][3, 2, 1])) < 1e-4 * max(1.0, abs(fd))


def test_energy_edge_cases():
    """Empty levels, a single conformation, lean mode, term subsets, offset torsions."""
    import grappa_oracle as orc
    from grappa_b200 import synthetic
    from grappa_b200.energy import Energy
    rng = np.random.default_rng(1)
    g = synthetic.make_molecule(rng, "small", n_confs=1, n_atoms=3)      # 2 bonds, 1 angle, no torsions
    assert g.num_nodes("n4") == 0 and g.num_nodes("n4_improper") == 0
    g.nodes["n2"].data["k"] = torch.tensor([300.0, 310.0]); g.nodes["n2"].data["eq"] = torch.tensor([1.1, 1.2])
    g.nodes["n3"].data["k"] = torch.tensor([80.0]); g.nodes["n3"].data["eq"] = torch.tensor([1.9])
    g.nodes["n4"].data["k"] = torch.zeros(0, 3); g.nodes["n4_improper"].data["k"] = torch.zeros(0, 3)
    idxs = {l: g.nodes[l].data["idxs"] for l in LEVELS}
    counts = {l: g.batch_num_nodes(l).tolist() for l in LEVELS}
    prm = {l: {n: g.nodes[l].data[n].double() for n in ("k", "eq") if n in g.nodes[l].data} for l in LEVELS}
    ref = orc.energy_forward(g.nodes["n1"].data["xyz"].double(), idxs, prm, counts)
    gd = Energy(write_tuple_terms=False)(g.to("cuda"))
    assert rel_err(gd.nodes["g"].data["energy"].cpu().numpy(), ref["energy"].detach().numpy()) < TOL_E
    assert rel_err(gd.nodes["n1"].data["gradient"].cpu().numpy(), ref["gradient"].numpy()) < TOL_E
    assert "x" not in gd.nodes["n2"].data
    # subset of terms
    gd2 = Energy(terms=["n2"], gradients=False)(g.to("cuda"))
    assert rel_err(gd2.nodes["g"].data["energy"].cpu().numpy(), ref["term_energy"]["n2"].numpy()) < TOL_E
    assert "gradient" not in gd2.nodes["n1"].data
    with pytest.raises(ValueError):
        Energy(terms=["n5"])(g.to("cuda"))
    # offset torsion on a molecule with torsions
    g = synthetic.make_molecule(rng, "peptide", n_confs=3, n_res=1)
    for l in LEVELS:
        T = g.num_nodes(l)
        g.nodes[l].data["k"] = torch.rand(T) * 100 if l in ("n2", "n3") else torch.randn(T, 3)
        if l in ("n2", "n3"):
            g.nodes[l].data["eq"] = torch.rand(T) + 1.0
    a = Energy(offset_torsion=False)(g.to("cuda")).nodes["g"].data["energy"].cpu()
    b = Energy(offset_torsion=True)(g.to("cuda")).nodes["g"].data["energy"].cpu()
    off = g.nodes["n4"].data["k"].abs().sum() + g.nodes["n4_improper"].data["k"].abs().sum()
    assert torch.allclose(b - a, off.expand_as(a), rtol=1e-5, atol=1e-3)


def test_energy_full_size_properties():
    """BASELINE config sizes (32 x 52 atoms x 50 conformations): size-independent properties.

    (i) translation + rotation invariance of energies, equivariance of forces; (ii) net force and net
    torque vanish per (molecule, conformation); (iii) conformation-permutation equivariance;
    (iv) tiled and global kernels agree."""
    from grappa_b200 import synthetic
    from grappa_b200.energy import Energy
    g = synthetic.peptide_batch(seed=0, batch_size=32, n_res=4, n_confs=50)
    gen = torch.Generator().manual_seed(2)
    for l in LEVELS:
        T = g.num_nodes(l)
        if l in ("n2", "n3"):
            g.nodes[l].data["k"] = 100 + 300 * torch.rand(T, generator=gen)
            g.nodes[l].data["eq"] = 1.2 + 0.6 * torch.rand(T, generator=gen)
        else:
            g.nodes[l].data["k"] = torch.randn(T, 3, generator=gen)
    gd = g.to("cuda")
    e1 = Energy(write_tuple_terms=False)
    out = e1(gd)
    E = out.nodes["g"].data["energy"].clone()
    F = out.nodes["n1"].data["gradient"].clone()
    B, C = E.shape
    assert (B, C) == (32, 50) and F.shape == (32 * 52, 50, 3)
    Fm = F.view(32, 52, 50, 3)
    X = gd.nodes["n1"].data["xyz"].view(32, 52, 50, 3)
    scale = F.abs().max()
    assert Fm.sum(1).abs().max() < 1e-4 * scale * 52                      # net force
    assert torch.cross(X, Fm, dim=-1).sum(1).abs().max() < 1e-3 * scale * 52   # net torque
    # rigid motion
    q, _ = torch.linalg.qr(torch.randn(3, 3, generator=gen))
    q = q.cuda()
    g2 = g.to("cuda")
    g2.nodes["n1"].data["xyz"] = gd.nodes["n1"].data["xyz"] @ q.T + torch.tensor([1.0, -2.0, 0.5], device="cuda")
    out2 = e1(g2)
    assert rel_err(out2.nodes["g"].data["energy"].cpu().numpy(), E.cpu().numpy()) < 5e-5
    assert rel_err(out2.nodes["n1"].data["gradient"].cpu().numpy(), (F @ q.T).cpu().numpy()) < 5e-5
    # conformation permutation
    perm = torch.randperm(50, generator=gen).cuda()
    g3 = g.to("cuda")
    g3.nodes["n1"].data["xyz"] = gd.nodes["n1"].data["xyz"][:, perm].contiguous()
    out3 = e1(g3)
    assert rel_err(out3.nodes["g"].data["energy"].cpu().numpy(), E[:, perm].cpu().numpy()) < 1e-6
    # kernel variants agree
    e2 = Energy(write_tuple_terms=False)
    e2.kernel_variant = 1
    out4 = e2(g.to("cuda"))
    assert rel_err(out4.nodes["g"].data["energy"].cpu().numpy(), E.cpu().numpy()) < 1e-5
    assert rel_err(out4.nodes["n1"].data["gradient"].cpu().numpy(), F.cpu().numpy()) < 1e-5


@pytest.mark.parametrize("variant", [0, 1, 3, 4, 5])
def test_energy_collinear_geometry_is_finite_and_matches_the_oracle(variant):
    """A linear molecule stored on an axis (CO2 / nitrile / alkyne type): |a x b| = 0 exactly.  The reference returns
    theta = atan2(0, c) = pi with a zero angle derivative (torch.norm's subgradient at 0); an rsqrt(0) = inf would turn
    energy, forces and -- through the loss -- every parameter into NaN.  Torsions over three collinear atoms are defined
    as phi = 0 with zero derivative (the reference masks that singularity with random noise)."""
    import grappa_oracle as orc
    from grappa_b200.energy import Energy
    from grappa_b200.graph import MolGraph
    C = 4
    # four atoms on the x axis, C conformations differing by a shift / scale along the axis (still exactly collinear)
    base = torch.tensor([[0.0, 0, 0], [1.25, 0, 0], [2.5, 0, 0], [3.5, 0, 0]])
    xyz = torch.stack([base * (1.0 + 0.125 * c) + torch.tensor([0.5 * c, 0, 0]) for c in range(C)], dim=1)   # (4, C, 3)
    src = torch.tensor([0, 1, 1, 2, 2, 3]); dst = torch.tensor([1, 0, 2, 1, 3, 2])
    g = MolGraph({"g": 1, "n1": 4, "n2": 3, "n3": 2, "n4": 1, "n4_improper": 0}, src, dst)
    g.nodes["n1"].data["xyz"] = xyz
    g.nodes["n2"].data["idxs"] = torch.tensor([[0, 1], [1, 2], [2, 3]])
    g.nodes["n3"].data["idxs"] = torch.tensor([[0, 1, 2], [1, 2, 3]])
    g.nodes["n4"].data["idxs"] = torch.tensor([[0, 1, 2, 3]])
    g.nodes["n4_improper"].data["idxs"] = torch.zeros((0, 4), dtype=torch.int64)
    g.nodes["n2"].data["k"] = torch.tensor([300.0, 310.0, 290.0]); g.nodes["n2"].data["eq"] = torch.tensor([1.1, 1.2, 1.0])
    g.nodes["n3"].data["k"] = torch.tensor([80.0, 70.0]); g.nodes["n3"].data["eq"] = torch.tensor([2.9, 3.0])
    g.nodes["n4"].data["k"] = torch.tensor([[0.5, -0.25, 0.125]])
    g.nodes["n4_improper"].data["k"] = torch.zeros(0, 3)
    # oracle on bonds + angles (its noise-free dihedral has a NaN gradient at the singularity, as the reference would)
    idxs = {l: g.nodes[l].data["idxs"] for l in ("n2", "n3")}
    counts = {l: g.batch_num_nodes(l).tolist() for l in ("n2", "n3")}
    prm = {l: {n: g.nodes[l].data[n].double() for n in ("k", "eq")} for l in ("n2", "n3")}
    ref = orc.energy_forward(xyz.double(), idxs, prm, counts, terms=("n2", "n3"))
    gd = g.to("cuda")
    for l in LEVELS:
        for n in ("k", "eq"):
            if n in gd.nodes[l].data:
                gd.nodes[l].data[n].requires_grad_(True)
    en = Energy()
    en.kernel_variant = variant
    out = en(gd)
    E = out.nodes["g"].data["energy"]
    F = out.nodes["n1"].data["gradient"]
    assert torch.isfinite(E).all() and torch.isfinite(F).all()
    e_tors = float(g.nodes["n4"].data["k"].sum())                  # phi = 0: sum_n k_n cos(0)
    assert rel_err((E.detach().cpu() - e_tors).numpy(), ref["energy"].detach().numpy()) < TOL_E
    assert rel_err(F.detach().cpu().numpy(), ref["gradient"].numpy()) < TOL_E
    assert torch.allclose(out.nodes["n3"].data["x"].cpu(), torch.full((2, C), float(np.pi)), atol=1e-6)
    # the double backward (K14) stays finite as well
    (E.sum() + (F * torch.randn_like(F)).sum()).backward()
    for l in LEVELS:
        for n in ("k", "eq"):
            t = gd.nodes[l].data.get(n)
            if t is not None and t.grad is not None:
                assert torch.isfinite(t.grad).all(), (l, n)
