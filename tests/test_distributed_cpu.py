"""N > 1 host logic on CPU: world_size-2 gloo processes exercise molecule sharding and the bucketed
gradient exchange of training.Trainer (flat gradient buffer, every bucket reduced exactly once)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out):
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, root)
    sys.path.insert(0, os.path.join(root, "oracle"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    import grappa_oracle as orc
    from grappa_b200 import models
    from grappa_b200.training import Trainer, init_distributed, shard_molecules
    r, _, w = init_distributed(backend="gloo")
    assert (r, w) == (rank, world)
    torch.manual_seed(0)
    model = models.model_from_config(orc.small_model_config())
    tr = Trainer(model, None, None, device="cpu")
    assert tr.distributed and tr.world == world and tr.comm_stream is None
    # gradients are garbage until the hook of their part fires: a part exchanged too early sums the garbage
    tr.fp.grad.fill_(1000.0)
    final = float(rank + 1)

    def finish(spans):
        for s, e in spans:
            tr.fp.grad[s:e] = final
    # fire the hooks in the order the backward pass does
    w_ = model.parameter_writer
    # rank 1's batch has no propers: that writer never runs there, its (zero) bucket is only marked ready -- and the
    # hooks of the other writers fire in a different order than on rank 0.  The collectives must still pair up.
    names = {id(mod): name for name, mod in tr.buckets}
    if rank == 1:
        finish([tr._bucket_spans["proper"]])
        tr._mark_unused("proper")
        fire = (w_.bond_writer, w_.improper_writer, w_.angle_writer)
    else:
        fire = (w_.improper_writer, w_.proper_writer, w_.angle_writer, w_.bond_writer)
    for mod in fire:
        # a writer's gradients become final part by part (symmetriser, transformer layers, then the rest)
        parts = tr._part_spans[names[id(mod)]]
        for k in range(len(mod.parts())):
            finish(parts[k])
            tr._on_stage_backward(("writer_part", (mod, k, ())))
        finish(parts[-1])
        tr._on_stage_backward(("writer", mod))
    for i in reversed(range(len(model.gnn.att_blocks))):
        finish([tr._block_spans[i]])
        tr._on_stage_backward(("gnn_block", i))
    (s0, e0), b0, b1 = tr._gnn_span, tr._block_spans[0][0], tr._block_spans[-1][1]
    finish([(s0, b0), (b1, e0)])          # pre_dense / post_dense: the GNN span outside its blocks
    tr._on_stage_backward(("gnn_rest", None))
    ok = bool(torch.all(tr.fp.grad == float(sum(range(1, world + 1)))))
    shard = list(shard_molecules(10, rank, world))
    out.put((rank, ok, shard))
    dist.barrier()
    dist.destroy_process_group()


def test_bucketed_allreduce_and_sharding_world2():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=240) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(ok for _, ok, _ in res), "a bucket was reduced twice or not at all"
    assert res[0][2] == [0, 2, 4, 6, 8] and res[1][2] == [1, 3, 5, 7, 9]


def test_gradient_buckets_tile_the_flat_buffer_exactly_once():
    """Every float of the flat gradient buffer belongs to exactly one exchanged span (writer parts, GNN blocks, the rest of
    the GNN) -- for the released architecture and for the narrow one, with and without convolution blocks."""
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, os.path.join(root, "oracle"))
    import grappa_oracle as orc
    from grappa_b200 import models
    from grappa_b200.training import Trainer
    for cfg in (orc.small_model_config(), dict(orc.small_model_config(), gnn_convolutions=1)):
        torch.manual_seed(0)
        model = models.model_from_config(cfg)
        tr = Trainer(model, None, None, device="cpu", distributed=False)
        spans = [sp for name in tr._part_spans for part in tr._part_spans[name] for sp in part]
        spans += list(tr._block_spans)
        s0, e0 = tr._gnn_span
        if tr._block_spans:
            spans += [(s0, tr._block_spans[0][0]), (tr._block_spans[-1][1], e0)]
        else:
            spans.append((s0, e0))
        cover = torch.zeros(tr.fp.total, dtype=torch.int32)
        for s, e in spans:
            assert 0 <= s <= e <= tr.fp.total and s % 4 == 0       # 16-byte aligned starts (peer_allreduce_kernel)
            cover[s:e] += 1
        assert bool((cover == 1).all()), "a gradient span is exchanged twice or never"
        order = tr._part_order()
        assert sorted(order) == sorted((n, k) for n in tr._part_spans for k in range(len(tr._part_spans[n])))
