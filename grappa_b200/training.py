"""Data-parallel training step around the drop-in modules (SURVEY.md sections 3.1, 8e).

The reference trains `Sequential(GrappaModel, Energy)` under PyTorch Lightning on ONE GPU
(training/trainrun.py:112-166, lightning_model.py:205-230: forward, MolwiseLoss, backward,
clip_grad_norm 10, Adam).  Lightning is neither available here nor on the hot path, so this module
provides the equivalent plain loop, B200-first:

  * parameters, gradients and Adam moments live in ONE flat fp32 buffer each (`FlatParams`); the
    modules' Parameters are views into it, so checkpoints / state_dict are unchanged
  * the backward kernels write weight gradients straight into the flat gradient buffer (no
    per-parameter accumulate kernels, no flatten / unflatten copies)
  * one process per GPU; gradients are averaged with bucketed all-reduces on a side stream that start as
    soon as a bucket's backward has finished (every writer in 5 parts, then the GNN blocks last-to-first),
    overlapping communication with the remaining backward kernels.  Single node: our own kernel over
    NVLink peer memory (`peer.PeerGradients`, csrc/peer_allreduce.cu); otherwise NCCL
  * global-norm clipping + Adam are two fused kernels over the flat buffers

Each rank draws its own batch (weak scaling); the global loss is the mean of the rank losses.
"""
from __future__ import annotations

import os
from typing import List, Optional, Tuple

import torch
import torch.distributed as dist

from . import models, ops


class FlatParams:
    """Re-homes all parameters of a module into one flat buffer (+ flat grads, Adam m / v)."""

    def __init__(self, module: torch.nn.Module, device, grad_factory=None):
        """`grad_factory(n_floats) -> fp32 tensor` supplies the gradient buffer (peer-mapped memory for the NVLink
        all-reduce, `peer.PeerGradients`); default: an ordinary device tensor."""
        self.module = module
        params = [p for p in module.parameters() if p.requires_grad]
        self.params = params
        sizes = [p.numel() for p in params]
        # 16-byte aligned slices so that TMA descriptors stay legal on the views
        offs, off = [], 0
        for n in sizes:
            offs.append(off)
            off += (n + 3) // 4 * 4
        self.offsets, self.total = offs, off
        self.flat = torch.zeros(off, device=device, dtype=torch.float32)
        self.grad = torch.zeros(off, device=device, dtype=torch.float32) if grad_factory is None else grad_factory(off)
        assert self.grad.numel() >= off and self.grad.dtype == torch.float32
        self.m = torch.zeros(off, device=device, dtype=torch.float32)
        self.v = torch.zeros(off, device=device, dtype=torch.float32)
        with torch.no_grad():
            for p, o in zip(params, offs):
                view = self.flat[o:o + p.numel()].view_as(p)
                view.copy_(p.data.to(device))
                p.data = view
                gv = self.grad[o:o + p.numel()].view_as(p)
                p.grad = gv
                p._gb_sink = gv          # backward kernels write here directly (models._StageFn)
        self._index = {id(p): i for i, p in enumerate(params)}

    def span(self, module: torch.nn.Module) -> Tuple[int, int]:
        """[start, end) of the flat slice covering all parameters of a sub-module (must be contiguous)."""
        idx = sorted(self._index[id(p)] for p in module.parameters() if id(p) in self._index)
        if not idx:
            return 0, 0
        assert idx == list(range(idx[0], idx[-1] + 1)), "sub-module parameters are not contiguous in the flat buffer"
        last = idx[-1]
        return self.offsets[idx[0]], self.offsets[last] + (self.params[last].numel() + 3) // 4 * 4


# timing experiments only (tools/sessions): run the data-parallel step WITHOUT exchanging gradients, to separate the cost of
# the collectives from the cost of running several processes side by side.  The ranks' parameters diverge.
_SKIP_ALLREDUCE = os.environ.get("GRAPPA_B200_SKIP_ALLREDUCE") == "1"


class _Captured:
    """One captured training step: static input graph + CUDA graph + static loss."""
    __slots__ = ("graph", "g_static", "pack", "loss", "keys")


class Trainer:
    """forward -> loss -> backward -> (bucketed all-reduce) -> clip + Adam, all on the GPU.

    `use_cuda_graph=True` (default on CUDA): the whole step -- ~550 kernel launches, the bucket
    all-reduces on the side stream, clip + Adam -- is captured once per batch *shape signature*
    (`PackedBatch.signature()` + conformation count) and replayed with a single launch.  The step count,
    the learning rate and the dropout RNG offset live in device memory so that replays stay correct:
    `counters[0]` = completed steps (Adam bias correction), `counters[1]` = offset added to every dropout
    seed; a `tick` kernel at the end of the step advances both.  The first step of a new signature runs
    eagerly (it also performs the one-off kernel attribute set-up), the second one is captured.
    """

    # per-batch inputs of the captured step: coordinates, labels, features and the loss's per-batch fields -- the
    # padded-conformation bookkeeping (`n_valid` / `is_dummy`, reference utils/dgl_utils.py:63-118) and the per-molecule
    # classical-parameter weights (`param_weight`, training/loss.py:72-76).  A field missing here would be deleted from
    # the static graph and the replayed step would optimise a different loss than the eager one.
    INPUT_KEYS = ("xyz", "energy_ref", "gradient_ref", "partial_charge", "n_valid", "is_dummy", "param_weight")

    def __init__(self, model: models.GrappaModel, energy, loss_fn, lr: float = 1.5e-5, clip: float = 10.0,
                 betas=(0.9, 0.999), eps: float = 1e-8, device="cuda", distributed: Optional[bool] = None,
                 use_cuda_graph: Optional[bool] = None, max_graphs: int = 8):
        self.model, self.energy, self.loss_fn = model, energy, loss_fn
        self.clip, self.betas, self.eps = clip, betas, eps
        self.device = torch.device(device)
        model.to(self.device)
        self.distributed = dist.is_available() and dist.is_initialized() if distributed is None else distributed
        self.world = dist.get_world_size() if self.distributed else 1
        # Single-node data parallelism on GPUs: gradients live in peer-mapped memory and are summed by our own NVLink
        # kernel (csrc/peer_allreduce.cu).  GRAPPA_B200_PEER_ALLREDUCE=0, more than 8 ranks or several nodes: NCCL.
        self.peer = None
        if (self.distributed and self.device.type == "cuda" and 1 < self.world <= 8
                and os.environ.get("GRAPPA_B200_PEER_ALLREDUCE", "1") != "0"
                and int(os.environ.get("LOCAL_WORLD_SIZE", self.world)) == self.world):
            from .peer import PeerGradients, PeerUnavailable
            holder = {}

            def factory(n):
                try:
                    holder["peer"] = PeerGradients(n, self.device)
                except PeerUnavailable as e:     # raised on all ranks alike: every rank falls back to NCCL
                    import warnings
                    warnings.warn(f"grappa_b200: {e}; gradients are exchanged with NCCL instead")
                    return torch.zeros(n, device=self.device, dtype=torch.float32)
                return holder["peer"].grad
            self.fp = FlatParams(model, self.device, grad_factory=factory)
            self.peer = holder.get("peer")
        else:
            self.fp = FlatParams(model, self.device)
        self.gnorm_sq = torch.zeros(1, device=self.device, dtype=torch.float32)
        self._norm_ws = torch.zeros(1024, device=self.device, dtype=torch.float32)
        self.counters = torch.zeros(2, device=self.device, dtype=torch.int64)
        self.lr_dev = torch.full((1,), float(lr), device=self.device, dtype=torch.float32)
        self._lr = float(lr)
        self._host_steps = 0
        if self.device.type == "cuda":
            ops.set_rng_offset(self.counters[1:2])
        self.use_cuda_graph = (self.device.type == "cuda") if use_cuda_graph is None else use_cuda_graph
        # every step (eager or replayed) is enqueued on the trainer's own stream: autograd binds its per-parameter
        # bookkeeping nodes to the stream they were first used on, and a stream capture may only depend on work of
        # the capturing stream -- so eager warm-up steps and the capture must share one non-default stream
        self.stream = torch.cuda.Stream(device=self.device) if self.device.type == "cuda" else None
        self.max_graphs = max_graphs
        self._captured = {}
        self._seen = {}
        self.last_graph = None
        # buckets in the order their gradients become available during backward
        w = model.parameter_writer
        g = model.gnn
        self.buckets: List[Tuple[str, torch.nn.Module]] = [
            ("improper", w.improper_writer), ("proper", w.proper_writer), ("angle", w.angle_writer), ("bond", w.bond_writer)]
        self._gnn_span = self.fp.span(g)
        self._bucket_spans = {name: self.fp.span(mod) for name, mod in self.buckets}
        # Every writer's bucket is exchanged in parts that become final one after the other during its backward pass
        # (symmetriser, transformer layers last to first, then the rest: projector / statistics), so the all-reduces of
        # the writers -- 54 % of the gradient bytes -- overlap the writers' own backward pass instead of queueing up
        # behind it (8 GPUs: the 13 whole-stage all-reduces ran back to back over the last 1.5 ms of the step and ended
        # 0.23 ms after the last compute kernel).  _part_spans[name][k] = list of flat spans; the last entry is the rest.
        self._part_spans = {}
        for name, mod in self.buckets:
            s0, e0 = self._bucket_spans[name]
            parts = [self.fp.span(m) for m in mod.parts()]
            rest, cur = [], s0
            for ps, pe in sorted(p for p in parts if p[1] > p[0]):
                if ps > cur:
                    rest.append((cur, ps))
                cur = max(cur, pe)
            if e0 > cur:
                rest.append((cur, e0))
            self._part_spans[name] = [[p] for p in parts] + [rest]
        self._block_spans = [self.fp.span(b) for b in g.att_blocks] if not g.no_convs else []
        self.comm_stream = torch.cuda.Stream(device=self.device) if self.distributed and self.device.type == "cuda" else None
        self._pending: List = []
        self._ready = {}           # (writer, part) -> events: gradients final, all-reduce not issued yet
        self._next_bucket = 0
        if self.distributed:
            models.set_backward_hook(self._on_stage_backward)

    # ---- learning rate / step count (device-resident) ------------------------------------------
    @property
    def lr(self) -> float:
        return self._lr

    @lr.setter
    def lr(self, value: float):
        if float(value) != self._lr:
            self._lr = float(value)
            self.lr_dev.fill_(self._lr)

    @property
    def step_count(self) -> int:
        return self._host_steps

    # ---- optimizer restart / checkpoint ---------------------------------------------------------
    def reset_optimizer(self, lr: Optional[float] = None):
        """A fresh Adam on the current parameters: moments and the bias-correction step count back to zero -- what the
        reference does when it replaces the optimizer at `param_loss_epochs` / after a restart
        (training/lightning_model.py:142-150).  In place, so captured steps stay valid; the dropout RNG offset keeps
        advancing."""
        self.fp.m.zero_()
        self.fp.v.zero_()
        self.counters[0:1].zero_()
        if lr is not None:
            self.lr = lr

    def state_dict(self) -> dict:
        """Everything a resumed run needs (reference: the Lightning checkpoint -- model + optimizer state + `lr`,
        training/lightning_model.py:300-304): parameters by name, Adam moments by name, step / RNG counters, lr."""
        names = {id(p): k for k, p in self.model.named_parameters()}
        sd = {"model": {k: v.detach().cpu().clone() for k, v in self.model.state_dict().items()},
              "exp_avg": {}, "exp_avg_sq": {}, "counters": self.counters.detach().cpu().clone(), "lr": self._lr,
              "host_steps": self._host_steps}
        for p, o in zip(self.fp.params, self.fp.offsets):
            n = p.numel()
            sd["exp_avg"][names[id(p)]] = self.fp.m[o:o + n].view_as(p).detach().cpu().clone()
            sd["exp_avg_sq"][names[id(p)]] = self.fp.v[o:o + n].view_as(p).detach().cpu().clone()
        return sd

    def load_state_dict(self, sd: dict):
        """Inverse of `state_dict`, written INTO the flat buffers (parameters stay views of them, captured steps stay
        valid).  Keyed by parameter name, so a checkpoint survives a different flattening order."""
        with torch.no_grad():
            self.model.load_state_dict(sd["model"])        # copy_ into the existing views
            names = {id(p): k for k, p in self.model.named_parameters()}
            for p, o in zip(self.fp.params, self.fp.offsets):
                n, k = p.numel(), names[id(p)]
                self.fp.m[o:o + n].view_as(p).copy_(sd["exp_avg"][k])
                self.fp.v[o:o + n].view_as(p).copy_(sd["exp_avg_sq"][k])
            self.counters.copy_(sd["counters"])
        self.lr = float(sd["lr"])
        self._host_steps = int(sd.get("host_steps", int(sd["counters"][0])))

    # ---- gradient exchange ---------------------------------------------------------------------
    def _launch_allreduce(self, start: int, end: int, events=None):
        """All-reduce grad[start:end] on the communication stream once `events` have completed (default: everything
        enqueued so far on the current stream)."""
        if end <= start or _SKIP_ALLREDUCE:
            return
        buf = self.fp.grad[start:end]
        if self.comm_stream is not None:
            if not events:
                ev = torch.cuda.Event()
                ev.record(torch.cuda.current_stream())
                events = (ev,)
            with torch.cuda.stream(self.comm_stream):
                for ev in events:
                    self.comm_stream.wait_event(ev)
                if self.peer is not None:
                    self.peer.allreduce(start, end - start)
                else:
                    dist.all_reduce(buf, op=dist.ReduceOp.SUM)
                done = torch.cuda.Event()
                done.record(self.comm_stream)
            self._pending.append(done)
        else:   # gloo / CPU tests
            dist.all_reduce(buf, op=dist.ReduceOp.SUM)

    def _on_stage_backward(self, tag):
        """Called from the backward pass when all gradients of a stage / GNN block are final."""
        kind, obj = tag
        if kind not in ("writer", "writer_part") and self.distributed:
            self._drain_writer_buckets(final=True)    # the GNN's backward starts after every writer's: nothing is left pending
        if kind == "writer_part":
            mod, k, events = obj
            for name, m in self.buckets:
                if m is mod:
                    self._ready[(name, k)] = events
            self._drain_writer_buckets()
        elif kind == "writer":
            for name, mod in self.buckets:
                if mod is obj:
                    self._ready[(name, len(self._part_spans[name]) - 1)] = ()
            self._drain_writer_buckets()
        elif kind == "gnn_block":
            self._launch_allreduce(*self._block_spans[obj])
        elif kind == "gnn_rest":
            # pre_dense / post_dense: everything of the GNN span outside the blocks
            s, e = self._gnn_span
            if self._block_spans:
                b0, b1 = self._block_spans[0][0], self._block_spans[-1][1]
                self._launch_allreduce(s, b0)
                self._launch_allreduce(b1, e)
            else:
                self._launch_allreduce(s, e)

    def _bucket_order(self) -> List[str]:
        """Writer buckets in the order autograd finishes them: the reverse of the order `WriteParameters.forward` ran the
        writers in (concurrent streams: proper, angle, bond, improper; one stream: bond, angle, proper, improper)."""
        from . import tape as T_
        if self.device.type == "cuda" and T_.concurrency():
            return ["improper", "bond", "angle", "proper"]
        return ["improper", "proper", "angle", "bond"]

    def _mark_unused(self, name: str):
        """The writer `name` does not run in this step (no tuples at its level on this rank): all its parts count as
        final and are exchanged at their usual positions in the order, like on the ranks that do run it."""
        for k in range(len(self._part_spans[name])):
            self._ready[(name, k)] = ()

    def _part_order(self):
        """(writer, part) pairs in the ONE order every rank issues them in: part-major (all symmetrisers, all last
        layers, ...), writers in the order autograd runs their backward passes on the host."""
        names = self._bucket_order()
        depth = max(len(self._part_spans[n]) for n in names)
        order = []
        for k in range(depth):
            for n in names:
                nk = len(self._part_spans[n])
                # a writer with fewer parts: its LAST part (the rest) keeps the last position
                kk = k if k < nk - 1 else (nk - 1 if k == depth - 1 else None)
                if kk is not None:
                    order.append((n, kk))
        return order

    def _drain_writer_buckets(self, final: bool = False):
        """Issue the writer-part all-reduces in ONE canonical order on every rank.  NCCL matches collectives by issue
        order, so a rank whose batch has no tuples at some level (that writer never runs, its bucket is zero) must
        issue that bucket at the same position as the ranks that do run it -- not before its forward pass.  Each
        all-reduce waits (on the device) only for the events of its own part."""
        order = self._part_order()
        while self._next_bucket < len(order) and (final or order[self._next_bucket] in self._ready):
            name, k = order[self._next_bucket]
            events = self._ready.get((name, k), ())
            for s, e in self._part_spans[name][k]:
                self._launch_allreduce(s, e, events)
            self._next_bucket += 1

    def _wait_comm(self):
        for ev in self._pending:
            torch.cuda.current_stream().wait_event(ev)
        self._pending = []

    # ---- one optimisation step -----------------------------------------------------------------
    def forward_backward(self, g) -> torch.Tensor:
        from .pack import get_pack
        pack = get_pack(g)
        self._ready, self._next_bucket = {}, 0
        for (name, _), lvl in zip(self.buckets, (3, 2, 1, 0)):
            if pack.n_tuples[lvl] == 0:      # writer unused by this batch: its gradient is zero, not stale
                s, e = self._bucket_spans[name]
                self.fp.grad[s:e].zero_()
                self._mark_unused(name)
        g = self.model(g)
        g = self.energy(g)
        loss = self.loss_fn(g)
        loss.backward()
        if self.distributed:
            self._drain_writer_buckets(final=True)
        # the outputs left on the graph (h, k, eq, energy, gradient ...) must not keep the autograd graph -- and with
        # it per-parameter autograd nodes bound to this step's streams -- alive beyond the step
        for nt in g.ntypes:
            d = g.nodes[nt].data
            for k in list(d.keys()):
                if d[k].grad_fn is not None:
                    d[k] = d[k].detach()
        return loss.detach()

    def optimizer_step(self):
        self._wait_comm()
        ops.sumsq_det(self.fp.grad, self.gnorm_sq, self._norm_ws)
        ops.adam_step_dev(self.fp.flat, self.fp.grad, self.fp.m, self.fp.v, self.lr_dev, self.betas[0], self.betas[1],
                          self.eps, self.counters[0:1], gnorm_sq=self.gnorm_sq, clip=self.clip, grad_scale=1.0 / self.world)
        ops.tick(self.counters)

    def _eager_step(self, g) -> torch.Tensor:
        loss = self.forward_backward(g)
        self.optimizer_step()
        return loss

    # ---- CUDA-graph path ------------------------------------------------------------------------
    def _input_keys(self, g):
        names = set(self.INPUT_KEYS) | set(self.model.gnn.in_feat_name)
        keys = []
        for nt in g.ntypes:
            for k in g.nodes[nt].data.keys():
                if k in names or k.endswith("_ref"):
                    keys.append((nt, k))
        return keys

    def _signature(self, g):
        from .pack import get_pack
        xyz = g.nodes["n1"].data.get("xyz")
        keys = tuple((nt, k, tuple(g.nodes[nt].data[k].shape)) for nt, k in self._input_keys(g))
        return (get_pack(g).signature(), None if xyz is None else tuple(xyz.shape), keys)

    def _refresh_inputs(self, cap: _Captured, g):
        from .pack import get_pack
        for nt, k in cap.keys:
            cap.g_static.nodes[nt].data[k].copy_(g.nodes[nt].data[k], non_blocking=True)
        cap.pack.copy_from(get_pack(g))

    def reset_graphs(self):
        """Drop every captured step (the next step of a known shape is captured again)."""
        self._captured.clear()
        self.last_graph = None

    def h2d_bytes(self, g) -> int:
        """Bytes `step(g)` copies host->device for a host-resident batch (inputs + index tables)."""
        from .pack import get_pack
        n = sum(g.nodes[nt].data[k].numel() * g.nodes[nt].data[k].element_size() for nt, k in self._input_keys(g))
        return n + get_pack(g).bytes

    def _capture(self, g) -> _Captured:
        from .pack import get_pack
        cap = _Captured()
        cap.keys = self._input_keys(g)
        gs = g.to(self.device)
        if g.device.type == "cuda":      # never alias the caller's tensors
            for nt, k in cap.keys:
                gs.nodes[nt].data[k] = g.nodes[nt].data[k].clone()
        for nt in gs.ntypes:             # stale outputs of an earlier eager pass are not inputs
            for k in list(gs.nodes[nt].data.keys()):
                if (nt, k) not in cap.keys and k != "idxs":
                    del gs.nodes[nt].data[k]
        cap.pack = get_pack(gs)
        if cap.pack is getattr(g, "_pack_cache", None):
            cap.pack = cap.pack.to(self.device)
            gs._pack_cache = cap.pack
        cap.g_static = gs
        torch.cuda.synchronize(self.device)
        cap.graph = torch.cuda.CUDAGraph()
        # Scratch buffers cached by earlier eager launches must not be baked into the graph: they live in the regular
        # allocator pool, would be freed by the drop below and unmapped by the empty_cache() of the next capture while this
        # graph still writes to them.  Dropping the cache first makes the capture allocate its scratch in its own pool.
        ops.drop_workspaces()
        with torch.cuda.graph(cap.graph, stream=self.stream, capture_error_mode="thread_local"):   # a loader thread may pin memory meanwhile
            cap.loss = self._eager_step(gs)
        ops.drop_workspaces()            # scratch allocated while capturing belongs to the graph's pool
        return cap

    def step(self, g) -> torch.Tensor:
        """One optimisation step on batch `g` (host or device resident); returns the (device) loss."""
        self._host_steps += 1
        if self.stream is None:
            self.last_graph = g
            return self._eager_step(g)
        caller = torch.cuda.current_stream(self.device)
        self.stream.wait_stream(caller)
        with torch.cuda.stream(self.stream):
            loss = self._step_on_stream(g)
        caller.wait_stream(self.stream)
        return loss

    def _step_on_stream(self, g) -> torch.Tensor:
        if not self.use_cuda_graph:
            if g.device.type != "cuda":
                g = g.to(self.device, non_blocking=True)
            self.last_graph = g
            return self._eager_step(g)
        sig = self._signature(g)
        cap = self._captured.get(sig)
        if cap is None:
            n = self._seen.get(sig, 0)
            self._seen[sig] = n + 1
            if n == 0 or len(self._captured) >= self.max_graphs:
                if g.device.type != "cuda":
                    g = g.to(self.device, non_blocking=True)
                self.last_graph = g
                return self._eager_step(g)
            cap = self._capture(g)
            self._captured[sig] = cap
        else:
            self._refresh_inputs(cap, g)
        cap.graph.replay()
        self.last_graph = cap.g_static
        return cap.loss


def init_distributed(backend: Optional[str] = None) -> Tuple[int, int, int]:
    """(rank, local_rank, world) from the torchrun environment; initialises the process group."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        if backend == "nccl":
            torch.cuda.set_device(local)
        dist.init_process_group(backend=backend, rank=rank, world_size=world)
    return rank, local, world


def shard_molecules(n_items: int, rank: int, world: int) -> range:
    """Inference / energy sweeps: molecule i -> rank i mod world, no communication."""
    return range(rank, n_items, world)
