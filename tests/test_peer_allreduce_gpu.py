"""NVLink peer-memory all-reduce (csrc/peer_allreduce.cu, grappa_b200/peer.py) on 2 GPUs: sums against the exact
expectation and against NCCL, repeated launches (epochs), odd spans; then two data-parallel training steps with the peer
kernel against the same steps with NCCL.  Needs >= 2 GPUs (skipped on the single-GPU test box; run by tools/sessions/m2d.sh)."""
import os
import socket

import pytest
import torch
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _leave(q):
    """End a worker without tearing NCCL down (destroy_process_group can block while captured graphs that contain
    collectives are alive): flush the result queue, meet the peer once more, exit."""
    import torch.distributed as dist
    q.close()
    q.join_thread()
    torch.cuda.synchronize()
    dist.barrier()
    torch.cuda.synchronize()
    os._exit(0)


def _worker(rank, world, port, q):
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, root)
    sys.path.insert(0, os.path.join(root, "oracle"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank),
                      LOCAL_WORLD_SIZE=str(world))
    import torch.distributed as dist
    from grappa_b200.training import init_distributed
    init_distributed()
    dev = torch.device("cuda", rank)
    from grappa_b200.peer import PeerGradients
    n = 1_000_003
    pg = PeerGradients(n, dev)
    ok = True
    for it, (start, count) in enumerate([(0, n), (4, 1), (1024, 65537), (12, n - 12), (0, 4), (400000, 300001)]):
        g = torch.Generator(device="cpu").manual_seed(100 * it)
        full = [torch.randn(pg.n, generator=g) for _ in range(world)]        # every rank can rebuild every rank's data
        pg.grad.copy_(full[rank].to(dev))
        torch.cuda.synchronize()
        dist.barrier()
        pg.allreduce(start, count)
        torch.cuda.synchronize()
        expect = full[rank].clone()
        acc = torch.zeros(count)
        for r in range(world):                                                 # rank order, like the kernel
            acc += full[r][start:start + count]
        # the kernel works on whole 16-byte groups: a span that ends inside a group reduces the rest of the group too
        end4 = min(pg.n, start + (count + 3) // 4 * 4)
        acc4 = torch.zeros(end4 - start)
        for r in range(world):
            acc4 += full[r][start:end4]
        expect[start:end4] = acc4
        ok = ok and torch.equal(pg.grad.cpu(), expect)
        dist.barrier()
    ok = ok and not pg.timed_out()
    # against NCCL on the same data
    g = torch.Generator(device="cpu").manual_seed(7 + rank)
    x = torch.randn(pg.n, generator=g).to(dev)
    pg.grad.copy_(x)
    y = x.clone()
    torch.cuda.synchronize(); dist.barrier()
    pg.allreduce(0, pg.n)
    dist.all_reduce(y)
    torch.cuda.synchronize()
    ok = ok and torch.equal(pg.grad, y)      # two ranks: a + b is the same in any order
    q.put((rank, bool(ok)))
    _leave(q)


def _train_worker(rank, world, port, q, use_peer):
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, root)
    sys.path.insert(0, os.path.join(root, "oracle"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank),
                      LOCAL_WORLD_SIZE=str(world), GRAPPA_B200_PEER_ALLREDUCE="1" if use_peer else "0")
    import torch.distributed as dist
    import grappa_oracle as orc
    from grappa_b200 import models, ops, synthetic
    from grappa_b200.energy import Energy
    from grappa_b200.loss import MolwiseLoss
    from grappa_b200.training import Trainer, init_distributed
    init_distributed()
    dev = torch.device("cuda", rank)
    ops.set_matmul_precision("fp32")
    cfg = dict(orc.small_model_config())
    for k in ("gnn_dropout_attention", "gnn_dropout_initial", "gnn_dropout_final", "parameter_dropout"):
        cfg[k] = 0.0
    model = models.model_from_config(cfg)
    model.load_state_dict(synthetic.deterministic_state_dict(model.state_dict(), seed=5))
    tr = Trainer(model.train(), Energy(write_tuple_terms=False), MolwiseLoss(proper_regularisation=1e-3, improper_regularisation=1e-3),
                 lr=1e-3, clip=10.0, device=dev, use_cuda_graph=True)
    assert (tr.peer is not None) == use_peer
    g = synthetic.peptide_batch(seed=20 + rank, batch_size=4, n_res=2, n_confs=6)      # every rank its own batch
    losses = [float(tr.step(g).item()) for _ in range(5)]                                # eager, eager, captured replays
    torch.cuda.synchronize()
    flat = tr.fp.flat.detach().cpu()
    gathered = [None] * world
    dist.all_gather_object(gathered, flat)
    same = all(torch.equal(gathered[0], t) for t in gathered)
    q.put((rank, losses, flat.numpy().copy(), same, False if tr.peer is None else tr.peer.timed_out()))
    _leave(q)


def _spawn(target, world, extra=()):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=target, args=(r, world, port, q) + tuple(extra), daemon=True) for r in range(world)]
    for p in procs:
        p.start()
    try:
        res = sorted((q.get(timeout=240) for _ in range(world)), key=lambda t: t[0])
        for p in procs:
            p.join(timeout=60)
            assert p.exitcode == 0
    finally:
        for p in procs:          # never leave a worker behind (a stuck rank would hold the GPU box until its time limit)
            if p.is_alive():
                p.kill()
    return res


@pytest.mark.skipif(not torch.cuda.is_available() or torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_peer_allreduce_sums_exactly_and_matches_nccl():
    res = _spawn(_worker, 2)
    assert all(ok for _, ok in res)


@pytest.mark.skipif(not torch.cuda.is_available() or torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_data_parallel_steps_peer_kernel_equals_nccl():
    import numpy as np
    peer = _spawn(_train_worker, 2, (True,))
    nccl = _spawn(_train_worker, 2, (False,))
    for (r, lp, fp_, same_p, timed_out), (_, ln, fn, same_n, _) in zip(peer, nccl):
        assert same_p and same_n and not timed_out          # all ranks hold identical parameters after the steps
        assert lp == ln, (lp, ln)                          # two ranks: the sum does not depend on the order
        assert np.array_equal(fp_, fn)
