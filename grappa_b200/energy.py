"""`Energy`: drop-in replacement for grappa.models.energy.Energy backed by kernels K13 / K14.

Mirror of reference src/grappa/models/energy.py:74-145 -- same constructor arguments
(`terms, suffix, offset_torsion, write_suffix, gradients`), same graph fields read
(`n1.xyz`, `n{2,3,4,4_improper}.{idxs,k,eq}`) and written:

    g.nodes['g'].data['energy'+ws]            (B, C)  differentiable w.r.t. k / eq
    g.nodes['g'].data['energy_'+term+ws]      (B, C)  detached
    g.nodes[term].data['energy'+ws]           (T, C)  per-tuple energies   (write_tuple_terms=True)
    g.nodes[term].data['unpooled_energy'+s]   alias of the line above (energy.py:66)
    g.nodes[term].data['x']                   (T, C)  internal coordinates (always recomputed; the
                                              reference caches torsion angles, internal_coordinates.py:77-81)
    g.nodes['n1'].data['gradient'+ws]         (N, C, 3) = +dE/dxyz, differentiable w.r.t. k / eq

Forces are analytic (no autograd over xyz); the backward pass of this module IS kernel K14, i.e. the
double backward the reference obtains with `create_graph=True` (energy.py:139).  `write_tuple_terms=
False` is the lean mode used for throughput runs (outputs energy + gradient only).
"""
from __future__ import annotations

import ctypes as C
from typing import List

import torch

from . import _lib
from .pack import LEVELS, get_pack

_LEVEL_ID = {l: i for i, l in enumerate(LEVELS)}


def _ptr(t):
    return 0 if t is None or t.numel() == 0 else t.data_ptr()


def _fill_common(a: _lib.EnergyArgs, pack, xyz, ks, eqs, n_per, level_mask, offset_torsion):
    a.xyz = xyz.data_ptr()
    a.n_atoms, a.n_confs, a.n_mols = pack.n_atoms, xyz.shape[1], pack.n_mols
    a.max_atoms_per_mol = pack.max_atoms_per_mol
    a.atom_off = pack.ptr("atom_off")
    for l in range(4):
        a.idx[l] = pack.ptr(f"idx{l}")
        a.tup_off[l] = pack.ptr(f"tup_off{l}")
        a.n_tuples[l] = pack.n_tuples[l] if (level_mask >> l) & 1 else 0
        a.k[l] = _ptr(ks[l])
    a.eq[0], a.eq[1] = _ptr(eqs[0]), _ptr(eqs[1])
    a.n_per[0], a.n_per[1] = n_per
    a.level_mask = level_mask
    a.offset_torsion = int(offset_torsion)
    if getattr(pack, "sched_groups", 0):
        a.sched_groups = pack.sched_groups
        for l in range(4):
            a.sched[l] = pack.ptr(f"sched{l}")
            a.round_off[l] = pack.ptr(f"round_off{l}")
            a.max_tuples_per_mol[l] = pack.max_tuples_per_mol[l]
            a.max_rounds_per_mol[l] = pack.max_rounds_per_mol[l]


class _EnergyFn(torch.autograd.Function):
    """(k2, eq2, k3, eq3, k4, k4i) -> (energy, gradient, term energies, x, tuple energies)."""

    @staticmethod
    def forward(ctx, pack, xyz, level_mask, offset_torsion, want_grad, want_tuple_terms, variant, k2, eq2, k3, eq3, k4, k4i):
        _lib.require_cuda(xyz)
        lib = _lib.lib()
        ks = [None if t is None else t.detach().contiguous().float() for t in (k2, k3, k4, k4i)]
        eqs = [None if t is None else t.detach().contiguous().float() for t in (eq2, eq3)]
        xyz = xyz.detach().contiguous().float()
        n_per = (ks[2].shape[1] if ks[2] is not None and ks[2].dim() == 2 else 0,
                 ks[3].shape[1] if ks[3] is not None and ks[3].dim() == 2 else 0)
        B, Cn, N = pack.n_mols, xyz.shape[1], pack.n_atoms
        dev = xyz.device
        a = _lib.EnergyArgs()
        _fill_common(a, pack, xyz, ks, eqs, n_per, level_mask, offset_torsion)
        energy = torch.empty((B, Cn), device=dev, dtype=torch.float32)
        terms = torch.zeros((4, B, Cn), device=dev, dtype=torch.float32)
        grad = torch.empty((N, Cn, 3), device=dev, dtype=torch.float32) if want_grad else None
        a.energy = energy.data_ptr()
        a.grad = _ptr(grad)
        xs, tes = [], []
        for l in range(4):
            a.term_energy[l] = terms[l].data_ptr()
            T = pack.n_tuples[l]
            if want_tuple_terms and (level_mask >> l) & 1:
                x = torch.empty((T, Cn), device=dev, dtype=torch.float32)
                te = torch.empty((T, Cn), device=dev, dtype=torch.float32)
                a.x[l], a.tuple_energy[l] = _ptr(x), _ptr(te)
            else:
                x = te = None
            xs.append(x)
            tes.append(te)
        _lib.check(lib.grappa_b200_energy_fwd(C.byref(a), variant, _lib.current_stream_ptr()), "energy_fwd")
        ctx.pack, ctx.level_mask, ctx.offset_torsion, ctx.n_per = pack, level_mask, offset_torsion, n_per
        ctx.save_for_backward(xyz, *[t if t is not None else torch.empty(0, device=dev) for t in ks + eqs])
        ctx.present = [t is not None for t in (k2, eq2, k3, eq3, k4, k4i)]
        outs = [energy, grad if grad is not None else torch.empty(0, device=dev), terms]
        for l in range(4):
            outs.append(xs[l] if xs[l] is not None else torch.empty(0, device=dev))
            outs.append(tes[l] if tes[l] is not None else torch.empty(0, device=dev))
        ctx.mark_non_differentiable(*outs[2:])
        return tuple(outs)

    @staticmethod
    def backward(ctx, g_energy, g_grad, *_unused):
        lib = _lib.lib()
        xyz, k2, k3, k4, k4i, eq2, eq3 = ctx.saved_tensors
        pack = ctx.pack
        ks = [k2, k3, k4, k4i]
        eqs = [eq2, eq3]
        ba = _lib.EnergyBwdArgs()
        _fill_common(ba.fwd, pack, xyz, [t if t.numel() else None for t in ks], [t if t.numel() else None for t in eqs],
                     ctx.n_per, ctx.level_mask, ctx.offset_torsion)
        ge = g_energy.contiguous().float() if g_energy is not None else None
        gg = g_grad.contiguous().float() if (g_grad is not None and g_grad.numel()) else None
        ba.g_energy, ba.g_grad = _ptr(ge), _ptr(gg)
        dks = [torch.zeros_like(t) for t in ks]
        deqs = [torch.zeros_like(t) for t in eqs]
        for l in range(4):
            ba.dk[l] = _ptr(dks[l]) if (ctx.level_mask >> l) & 1 else 0
        ba.deq[0] = _ptr(deqs[0]) if ctx.level_mask & 1 else 0
        ba.deq[1] = _ptr(deqs[1]) if ctx.level_mask & 2 else 0
        _lib.check(lib.grappa_b200_energy_bwd(C.byref(ba), _lib.current_stream_ptr()), "energy_bwd")
        pres = ctx.present
        grads = [dks[0] if pres[0] else None, deqs[0] if pres[1] else None, dks[1] if pres[2] else None,
                 deqs[1] if pres[3] else None, dks[2] if pres[4] else None, dks[3] if pres[5] else None]
        return (None, None, None, None, None, None, None, *grads)


def energy_and_gradient(pack, xyz, params, level_mask=0b1111, offset_torsion=False, gradients=True,
                        write_tuple_terms=False, variant=0):
    """Functional entry: params = {'n2': {'k','eq'}, 'n3': {...}, 'n4': {'k'}, 'n4_improper': {'k'}}."""
    def get(lvl, name):
        d = params.get(lvl)
        return None if d is None else d.get(name)
    return _EnergyFn.apply(pack, xyz, level_mask, offset_torsion, gradients, write_tuple_terms, variant,
                           get("n2", "k"), get("n2", "eq"), get("n3", "k"), get("n3", "eq"),
                           get("n4", "k"), get("n4_improper", "k"))


class Energy(torch.nn.Module):
    """Writes the bonded MM energy (and its gradient w.r.t. positions) of all conformations into the graph."""

    def __init__(self, terms: List[str] = ["n2", "n3", "n4", "n4_improper"], suffix: str = "",
                 offset_torsion: bool = False, write_suffix=None, gradients: bool = True,
                 write_tuple_terms: bool = True):
        super().__init__()
        if not isinstance(terms, list):
            raise ValueError("terms must be a list")
        self.offset_torsion = offset_torsion
        self.suffix = suffix
        self.write_suffix = write_suffix if write_suffix is not None else suffix
        self.terms = terms
        self.gradients = gradients
        self.write_tuple_terms = write_tuple_terms
        self.kernel_variant = 0

    def forward(self, g):
        if "xyz" not in g.nodes["n1"].data.keys():
            raise ValueError("xyz coordinates must be stored in g.nodes['n1'].data['xyz']")
        xyz = g.nodes["n1"].data["xyz"]
        _lib.require_cuda(xyz)
        mask = 0
        params = {}
        for term in self.terms:
            if term not in g.ntypes:
                raise ValueError(f"term {term} not in g.ntypes")
            if term not in _LEVEL_ID:
                raise ValueError(f"unknown term {term}")
            d = g.nodes[term].data
            if "k" + self.suffix not in d.keys():
                raise RuntimeError(f"{term} has no k{self.suffix} attribute")
            params[term] = {"k": d["k" + self.suffix]}
            if term in ("n2", "n3"):
                params[term]["eq"] = d["eq" + self.suffix]
            mask |= 1 << _LEVEL_ID[term]
        pack = get_pack(g)
        outs = energy_and_gradient(pack, xyz, params, mask, self.offset_torsion, self.gradients,
                                   self.write_tuple_terms, self.kernel_variant)
        energy, grad, terms = outs[0], outs[1], outs[2]
        ws = self.write_suffix
        for term in self.terms:
            l = _LEVEL_ID[term]
            g.nodes["g"].data["energy_" + term + ws] = terms[l]
            if self.write_tuple_terms:
                g.nodes[term].data["x"] = outs[3 + 2 * l]
                g.nodes[term].data["energy" + ws] = outs[4 + 2 * l]
                g.nodes[term].data["unpooled_energy" + self.suffix] = outs[4 + 2 * l]
        g.nodes["g"].data["energy" + ws] = energy
        if self.gradients:
            g.nodes["n1"].data["gradient" + ws] = grad
        return g
