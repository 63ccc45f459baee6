"""One K13 launch configuration for ncu: python tools/energy_one.py [n_mols n_confs variant [full]]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from grappa_b200 import graph as gbg, synthetic
from grappa_b200.energy import Energy
n_mols = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
n_confs = int(sys.argv[2]) if len(sys.argv) > 2 else 100
variant = int(sys.argv[3]) if len(sys.argv) > 3 else 0
dev = torch.device("cuda")
ge = synthetic.peptide_batch(seed=7, batch_size=8, n_res=4, n_confs=n_confs)
ge = gbg.batch([ge] * (n_mols // 8)).to(dev)
gen = torch.Generator().manual_seed(1)
for l in ("n2", "n3", "n4", "n4_improper"):
    T = ge.num_nodes(l)
    if l in ("n2", "n3"):
        ge.nodes[l].data["k"] = (100 + 300 * torch.rand(T, generator=gen)).to(dev)
        ge.nodes[l].data["eq"] = (1.2 + 0.6 * torch.rand(T, generator=gen)).to(dev)
    else:
        ge.nodes[l].data["k"] = torch.randn(T, 3, generator=gen).to(dev)
en = Energy(write_tuple_terms=(len(sys.argv) > 4 and sys.argv[4] == "full"))
en.kernel_variant = variant
with torch.no_grad():
    for _ in range(3):
        en(ge)
torch.cuda.synchronize()
