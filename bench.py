#!/usr/bin/env python
"""Headline benchmark of the Grappa hot path on B200 (contract: see the task prompt / DESIGN.md).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--precision bf16x3|tf32|fp32|policy]

Workload = BASELINE.json configs[1]: one training step on a batch of 32 synthetic peptides
(ACE-(ALA)4-NME, 52 atoms, 50 conformations each) per GPU: grappa-1.2 GNN + 4 writers + MM
energy/forces + molecule-wise energy+force loss, backward, global-norm clip, Adam.  Dropout is ON
(train mode), nothing is cached between steps.  N > 1: one process per GPU (torchrun), each rank its
own batch (weak scaling), bucketed gradient all-reduce overlapped with backward -- our own kernel over NVLink peer
memory (csrc/peer_allreduce.cu; GRAPPA_B200_PEER_ALLREDUCE=0 selects NCCL), captured into the step's CUDA graph.

Printed JSON line (rank 0): metric = training molecules/s (whole job); `e2e` = same through the public
API from pinned HOST buffers (H2D of the batch + D2H of the loss inside the timed region); `e2e_loader` =
the same with a FRESH batch collated by dataset.PrefetchLoader every step; `roofline` for the dominant kernel
family (GEMMs -> tensor pipe, burst peak: the family is replayed alone); `gather_kernels` = HBM fractions of the
gather-bound kernels north_star names (edge attention, tuple gather, LayerNorm); `energy_eval` = the second half of
BASELINE's metric (conformation energy+force evaluations/s of kernel K13 on configs[3]: 1k molecules per GPU x
{100, 1000} conformations, lean and full-contract outputs, molecules sharded over the ranks, HBM roofline);
`cpu_baseline` = the reference's CPU path on the host cores, same workload.

`--impl reference`: the reference's own implementation on the host CPU -- the UNMODIFIED reference package
(baseline/_ref, installed with pip from /root/reference; DGL replaced by oracle/dgl_shim because no DGL build exists
for this image) when present, else the oracle port.  It loads nothing of grappa_b200's native code.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOAD = "train_step_32xACE-ALA4-NME_52atoms_50confs"
METRIC = "training molecules/sec (grappa-1.2 fwd + energy/force loss + bwd + Adam); conformation energy+force evals/sec in energy_eval"


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return {"hbm_gbs": d["hbm_gbs"], "tflops_burst": d["bf16_tflops"], "tflops_sustained": d["bf16_tflops_sustained"],
                "source": "measured (MEASURED_PEAKS.json)"}
    return {"hbm_gbs": 6650.0, "tflops_burst": 1590.0, "tflops_sustained": 1400.0, "source": "fallback (B200_PROFILING.md)"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""

    def __init__(self, index: int):
        self.index, self.rows, self.proc = index, [], None

    def __enter__(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={q}", "--format=csv,noheader,nounits",
                                          "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def __exit__(self, *exc):
        if self.proc is not None:
            time.sleep(0.15)
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()

    def summary(self):
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
            except Exception:
                continue
            for n, v in zip(names, r[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": max(mx), "reasons": sorted(reasons), "samples": len(sm)}


# --------------------------------------------------------------------------------------------------
# reference arm: the reference's own CPU implementation of the path.  No native code of grappa_b200 is loaded here.
# --------------------------------------------------------------------------------------------------
def _reference_src():
    for cand in (os.path.join(ROOT, "baseline", "_ref"), "/root/reference/src"):
        if os.path.isdir(os.path.join(cand, "grappa", "models")):
            return cand
    return None


def _bench_batch_shim(batch_size: int, seed: int = 0):
    """Batch of `batch_size` ACE-(ALA)4-NME molecules (52 atoms, 50 conformations) as a heterograph of the DGL shim, from
    the committed single-molecule inputs (tests/golden/bench_peptide52.npz: topology, tuples, features, one embedding);
    every copy gets its own conformations / labels.  Plain numpy + torch: nothing of grappa_b200 is imported."""
    import numpy as np
    import torch
    shim = os.path.join(ROOT, "oracle", "dgl_shim")
    if shim not in sys.path:
        sys.path.insert(0, shim)
    import dgl
    z = np.load(os.path.join(ROOT, "tests", "golden", "bench_peptide52.npz"))
    rng = np.random.default_rng(seed)
    ntypes = ("g", "n1", "n2", "n3", "n4", "n4_improper")
    graphs = []
    for _ in range(batch_size):
        data = {("n1", "n1_edge", "n1"): (torch.from_numpy(z["in.src"]).long(), torch.from_numpy(z["in.dst"]).long())}
        num = {}
        for nt in ntypes:
            n = int(z[f"in.count.{nt}"][0])
            num[nt] = n
            if nt != "n1":
                data[(nt, f"{nt}_edge", nt)] = (torch.arange(n), torch.arange(n))
        g = dgl.heterograph(data, num)
        for k in z.files:
            if k.startswith("in.count.") or k in ("in.src", "in.dst"):
                continue
            nt, name = k[3:].split(".", 1)
            v = z[k]
            if name == "xyz":
                v = v + rng.normal(0.0, 0.05, size=v.shape).astype(np.float32)
            elif name == "gradient_ref":
                v = rng.normal(0.0, 10.0, size=v.shape).astype(np.float32)
            elif name == "energy_ref":
                v = rng.normal(0.0, 3.0, size=v.shape).astype(np.float32)
                v = v - v.mean(axis=1, keepdims=True)
            elif name == "partial_charge":
                v = np.clip(rng.normal(0.0, 0.3, size=v.shape), -1, 1).astype(np.float32)
            g.nodes[nt].data[name] = torch.from_numpy(np.ascontiguousarray(v))
        graphs.append(g)
    return dgl.batch(graphs)


def cpu_reference_step(steps: int, warmup: int, batch_size: int = 32):
    """One training step of the reference on the host CPU, all cores: forward (GrappaModel + Energy), MolwiseLoss, backward,
    clip_grad_norm_(10) + Adam (training/lightning_model.py:205-230,297-299; config.py:106), dropout on (train mode) --
    the reference as shipped.  Real reference when installed (baseline/_ref), else the oracle port (fwd + loss + bwd)."""
    import torch
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    g = _bench_batch_shim(batch_size)
    src = _reference_src()
    if src is not None:
        os.environ["GRAPPA_REFERENCE_SRC"] = src
        import ref_import
        ref_import.REFERENCE_SRC = src
        ns = ref_import.import_reference()
        import grappa_oracle as orc
        cfg = orc.grappa_1_2_model_config()
        torch.manual_seed(0)
        model = ns.deploy.model_from_config(dict(cfg), param_statistics=ns.graph_utils.get_default_statistics()).train()
        full = torch.nn.Sequential(model, ns.energy.Energy())
        loss_fn = ns.loss.MolwiseLoss(gradient_weight=0.8, energy_weight=1.0, param_weight=0.0, proper_regularisation=1e-3,
                                      improper_regularisation=1e-3)
        opt = torch.optim.Adam(model.parameters(), lr=1.5e-5)
        import copy

        def one():
            dg = full(copy.deepcopy(g))
            loss = loss_fn(dg)
            opt.zero_grad()
            loss.backward()
            torch.nn.utils.clip_grad_norm_(model.parameters(), 10.0)
            opt.step()
            return float(loss)
        kind = "reference"
        what = (f"UNMODIFIED reference package ({os.path.relpath(src, ROOT) if src.startswith(ROOT) else src}: grappa.models + "
                "Energy + training.loss.MolwiseLoss + clip + Adam, train mode), DGL calls served by oracle/dgl_shim (no DGL "
                "build exists for this image)")
    else:
        import grappa_oracle as orc
        from grappa_b200 import models      # parameter container only (python; the native library is not loaded)
        cfg = orc.grappa_1_2_model_config()
        torch.manual_seed(0)
        template = models.model_from_config(cfg)
        sd = {k: v.detach().clone().requires_grad_(v.is_floating_point() and v.dim() > 0 and "permutation" not in k
                                                   and "positional" not in k and "k_mean" not in k and "k_std" not in k
                                                   and "to_k" not in k and "to_eq" not in k)
              for k, v in template.state_dict().items()}
        leaves = [v for v in sd.values() if v.requires_grad]

        def one():
            h, params, en = orc.path_forward(sd, g, cfg, create_graph=True)
            loss = orc.molwise_loss(en, params, g)
            torch.autograd.grad(loss, leaves, allow_unused=True)
            return float(loss)
        kind = "port"
        what = "oracle/grappa_oracle.py (torch CPU fp32 restatement of the reference; fwd + loss + bwd, eval-mode dropout)"
    for _ in range(warmup):
        one()
    t0 = time.perf_counter()
    for _ in range(steps):
        one()
    dt = (time.perf_counter() - t0) / max(steps, 1)
    return {"value": batch_size / dt, "unit": "molecules/s", "cores": cores, "kind": kind,
            "sample": f"{steps} full steps of one {batch_size}-molecule batch after {warmup} warm-up: {what}", "s_per_step": dt}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    steps = max(1, min(args.steps, 3))
    warm = max(1, min(args.warmup, 1))
    r = cpu_reference_step(steps, warm)
    line = {"metric": METRIC, "value": r["value"], "unit": "molecules/s", "n_gpus": args.gpus, "steps": steps, "warmup": warm,
            "ms_per_step": r["s_per_step"] * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic", "impl": "reference",
            "config": {"workload": WORKLOAD, "note": "the reference's own CPU implementation of the path on all host cores; "
                       "steps are bounded (<= 3) so that the run ends within minutes"},
            "cpu_baseline": {k: r[k] for k in ("value", "unit", "cores", "kind", "sample")},
            "e2e": {"value": r["value"], "unit": "molecules/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--precision", default=None, help="GEMM arithmetic: bf16x3 (default, grappa_b200.ops.BENCH_PRECISION) | tf32 | fp32 "
                    "| a per-family policy such as 'fwd=bf16x3,dgrad=bf16x3,wgrad=tf32'")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip energy_eval / inference_eval / gather_kernels / e2e_loader")
    ap.add_argument("--no-cuda-graph", action="store_true", help="launch every kernel from the host instead of replaying a captured step")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist
    from grappa_b200 import _lib, models, ops, synthetic
    from grappa_b200.energy import Energy
    from grappa_b200.loss import MolwiseLoss
    from grappa_b200.pack import get_pack
    from grappa_b200.training import Trainer, init_distributed

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: grappa_b200 has no CPU path (use --impl reference for the CPU arm)")
    rank, local, world = init_distributed()
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    peaks = load_peaks()
    W = max(args.warmup, 3)
    K = args.steps
    if args.precision is None:
        args.precision = ops.BENCH_PRECISION
    ops.set_matmul_precision(args.precision)

    # ---- model / data ---------------------------------------------------------------------------
    torch.manual_seed(0)
    cfg = models.grappa_1_2_model_config()
    model = models.model_from_config(cfg).train()
    trainer = Trainer(model, Energy(write_tuple_terms=False), MolwiseLoss(gradient_weight=0.8, energy_weight=1.0,
                      param_weight=0.0, proper_regularisation=1e-3, improper_regularisation=1e-3), lr=1.5e-5, clip=10.0, device=dev)
    B = 32
    g_host = synthetic.peptide_batch(seed=100 + rank, batch_size=B, n_res=4, n_confs=50)
    get_pack(g_host)                    # host-side index tables (what a data-loader worker prepares)
    g_host = g_host.pin_memory()
    h2d = trainer.h2d_bytes(g_host)
    n_atoms = g_host.num_nodes("n1")

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def timed(fn, steps):
        sync_all()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        sync_all()
        return ms.item()

    def timed_local(fn, steps):
        """CUDA-event time of `steps` calls on this rank only (no collective): per-kernel micro-measurements."""
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1)

    # ---- (1) device-resident throughput ----------------------------------------------------------
    g_dev = g_host.to(dev)

    def step_resident():
        trainer.step(g_dev)

    # launches of one step, counted on an eager (un-captured) step: graph replays re-issue exactly these kernels
    trainer.use_cuda_graph = False
    step_resident()
    l0 = _lib.launch_count()
    step_resident()
    launches_per_step = _lib.launch_count() - l0
    trainer.use_cuda_graph = not args.no_cuda_graph
    for _ in range(W):
        step_resident()
    with ClockSampler(local) as clk:
        ms = timed(step_resident, K)
    launches = launches_per_step * K
    value = world * B * K / (ms * 1e-3)

    # multi-rank correctness evidence: after the timed steps every rank must hold bit-identical parameters (same initial
    # weights, summed gradients, the same deterministic clip + Adam on every rank)
    param_sync = None
    if world > 1:
        flat = trainer.fp.flat
        cks = torch.stack([flat.double().sum(), flat.double().abs().sum(), (flat.double() * torch.arange(
            flat.numel(), device=dev, dtype=torch.float64).remainder(97.0)).sum()])
        lo, hi = cks.clone(), cks.clone()
        dist.all_reduce(lo, op=dist.ReduceOp.MIN)
        dist.all_reduce(hi, op=dist.ReduceOp.MAX)
        param_sync = {"max_cross_rank_checksum_difference": float((hi - lo).abs().max().item()),
                      "identical_on_all_ranks": bool((hi - lo).abs().max().item() == 0.0),
                      "checksums": "sum, sum|.|, position-weighted sum of the flat fp32 parameter buffer (fp64), MIN / MAX over ranks",
                      "steps_before_check": int(trainer.step_count),
                      "gradient_exchange": ("peer_allreduce_kernel over NVLink peer memory (csrc/peer_allreduce.cu), "
                                            f"{trainer.peer.ctas} CTAs per launch" if trainer.peer is not None
                                            else "ncclAllReduce (torch.distributed)"),
                      "peer_barrier_timed_out": (bool(trainer.peer.timed_out()) if trainer.peer is not None else None)}

    # ---- (2) end to end from pinned host memory --------------------------------------------------
    def step_e2e():
        # public API with HOST buffers: the trainer copies the batch (features, xyz, labels, index tables) from pinned
        # host memory into its static device buffers (H2D), runs the step, and the loss is read back (D2H)
        loss = trainer.step(g_host)
        return loss.item()

    for _ in range(3):
        step_e2e()
    ms_e2e = timed(step_e2e, K)
    e2e_value = world * B * K / (ms_e2e * 1e-3)

    # ---- (2b) end to end with a FRESH batch per step from the packed dataset through the prefetching loader ---------
    e2e_loader = None
    if not args.no_extras:
        try:
            import numpy as np
            from grappa_b200 import dataset
            rng = np.random.default_rng(1000 + rank)
            mols = [synthetic.make_molecule(rng, "peptide", n_confs=50, n_res=4) for _ in range(64)]
            ds = dataset.PackedDataset.from_graphs(mols)
            n_batches = K + 4
            order = [rng.permutation(64)[:B].tolist() for _ in range(n_batches)]
            loader = iter(dataset.PrefetchLoader(ds, order, conf_strategy=50, seed=rank, depth=3, workers=2))
            for _ in range(4):
                trainer.step(next(loader)).item()

            def step_loader():
                return trainer.step(next(loader)).item()
            ms_ld = timed(step_loader, K)
            e2e_loader = {"value": world * B * K / (ms_ld * 1e-3), "unit": "molecules/s", "ms_per_step": ms_ld / K,
                          "what": "every step takes a NEW batch: 32 molecules drawn from a 64-molecule PackedDataset, collated "
                                  "(index offsets, conformation selection, conflict-free schedule, index tables) and pinned by "
                                  "dataset.PrefetchLoader (2 worker threads, depth 3), H2D + step + loss read-back"}
            del loader
            # the same stream of batches assembled ON THE DEVICE: the packed dataset and every molecule's index tables are
            # resident in HBM, a batch costs ~25 KB of host tables, one H2D copy of them and one collate kernel
            try:
                dd = dataset.DeviceDataset(ds, dev)
                state = {"i": 0}
                rng_d = np.random.default_rng(2000 + rank)

                def next_batch():
                    idx = order[state["i"] % n_batches]
                    state["i"] += 1
                    return dd.collate(idx, 50, rng_d)
                state["g"] = next_batch()

                def step_device_collate():
                    # software pipeline: the step on batch k is enqueued, then the host builds the job table of batch k + 1
                    # (its collate kernel queues up behind the step), then the loss of batch k is read back
                    loss = trainer.step(state["g"])
                    state["g"] = next_batch()
                    return loss.item()
                for _ in range(4):
                    step_device_collate()
                ms_dc = timed(step_device_collate, K)
                e2e_loader["device_collate"] = {
                    "value": world * B * K / (ms_dc * 1e-3), "unit": "molecules/s", "ms_per_step": ms_dc / K,
                    "what": "every step takes a NEW batch assembled by dataset.DeviceDataset.collate (kernel grappa_b200_collate) from "
                            "the HBM-resident dataset: host work = sampling + a ~25 KB job table (built for batch k + 1 while step k runs), no host gathers, "
                            "no H2D of features; loss of every step read back"}
            except Exception as e:  # pragma: no cover
                e2e_loader["device_collate"] = {"error": repr(e)}
        except Exception as e:  # pragma: no cover
            e2e_loader = {"error": repr(e)}

    # ---- (3) roofline of the dominant kernel family (GEMMs) ---------------------------------------------------------
    # Every GEMM call of one step is recorded (the exact gb_gemm_args, tensors kept alive) during an eager step, then
    # ALL of them are re-issued back to back on one stream as a captured graph and that graph is timed with CUDA events:
    # GEMM kernels only, nothing in between, no per-launch event overhead.  Inputs are colder than inside the step (the
    # producers' outputs are no longer in L2), so the figure is conservative.  The replay is a few ms of isolated tensor
    # work -> the BURST peak is the denominator.
    from grappa_b200 import tape as gb_tape
    gb_tape.set_concurrency(False)
    trainer.reset_graphs()
    trainer.use_cuda_graph = False
    rec = []
    ops.set_gemm_recorder(rec)
    step_resident()
    ops.set_gemm_recorder(None)
    torch.cuda.synchronize()
    l0 = _lib.launch_count()
    gemm_flops = ops.replay_gemms(rec)
    gemm_launches = _lib.launch_count() - l0
    torch.cuda.synchronize()
    gemm_graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(gemm_graph):
        ops.replay_gemms(rec)
    for _ in range(3):
        gemm_graph.replay()
    n_prof_steps = 10
    gemm_ms = timed(gemm_graph.replay, n_prof_steps)
    del gemm_graph
    gemm_calls = sum(n for _, _, n, _ in rec)
    by_prec = {}
    for kind, arr, n, _ in rec:
        for i in range(n):
            g_ = arr if kind == "single" else arr[i]
            nm = {0: "fp32", 1: "tf32", 2: "tf32", 3: "bf16x3"}[int(g_.precision)]
            by_prec[nm] = by_prec.get(nm, 0.0) + 2.0 * g_.M * g_.N * g_.K
    if os.environ.get("GRAPPA_B200_GEMM_TABLE") and rank == 0:
        # tuning aid: the recorded shapes
        agg = {}
        for kind, arr, n, _ in rec:
            for i in range(n):
                g_ = arr if kind == "single" else arr[i]
                k_ = (kind, g_.M, g_.N, g_.K, int(g_.trans_a), int(g_.trans_b))
                agg[k_] = agg.get(k_, 0) + 1
        with open(os.environ["GRAPPA_B200_GEMM_TABLE"], "w") as fh:
            for k_, c in sorted(agg.items(), key=lambda kv: -kv[1] * kv[0][1] * kv[0][2] * kv[0][3]):
                fh.write(f"{k_[0]:8s} M={k_[1]:6d} N={k_[2]:5d} K={k_[3]:6d} ta={k_[4]} tb={k_[5]}  x{c:3d}  "
                         f"{2.0 * c * k_[1] * k_[2] * k_[3] / 1e9:8.2f} GFLOP\n")
    rec.clear()
    trainer.use_cuda_graph = not args.no_cuda_graph
    trainer.reset_graphs()
    step_resident()
    ms_serial = timed(step_resident, 5) / 5
    gemm_flops *= n_prof_steps
    gb_tape.set_concurrency(True)
    trainer.reset_graphs()
    achieved_tf = gemm_flops / (gemm_ms * 1e-3) / 1e12 if gemm_ms > 0 else 0.0
    prec = ops.get_matmul_precision()
    tensor_path = prec != "fp32"
    dtype = {"fp32": "f32", "tf32": "tf32 (tcgen05 kind::tf32, fp32 accumulate)",
             "bf16x3": "bf16x3 (tcgen05 kind::f16 on in-kernel bf16 hi/lo splits of the fp32 operands: hi*hi + lo*hi + hi*lo, "
                       "fp32 accumulate in TMEM)"}.get(prec, prec)
    # executed tensor-pipe FLOPs per algorithmic FLOP: bf16x3 issues three bf16 MMAs per product, tf32 runs at half the bf16 rate
    mma_factor = {"bf16x3": 3.0, "tf32": 2.0, "fp32": 0.0}
    tot_alg = sum(by_prec.values()) or 1.0
    exec_equiv = sum(mma_factor.get(k, 0.0) * v for k, v in by_prec.items()) / tot_alg
    roofline = {"bound": "tensor", "kernel": "gemm_tc_kernel (tcgen05 + TMA)" if tensor_path else "sgemm_kernel (fp32 FFMA)",
                "achieved": achieved_tf, "peak": peaks["tflops_burst"], "unit": "TFLOP/s",
                "frac": achieved_tf / peaks["tflops_burst"],
                "traffic": 67.4e6 if prec == "bf16x3" else None,
                "traffic_source": "dram__bytes_read.sum + dram__bytes_write.sum = 33.6 + 33.7 MB for ONE launch of the step's largest "
                                  "token GEMM (proper in_proj, 14848 x 1536 x 512, bf16x3; ncu --set full, profiles/r2_gemm_x3_ncu_raw.csv); "
                                  "algorithmic bytes of that launch 124.7 MB (A 30.4 + B 3.1 + C 91.2): no operand is re-read from HBM, "
                                  "most of C is still in the 126 MB L2 when the kernel ends",
                "algorithmic_flops_by_arithmetic": {k: v for k, v in by_prec.items()},
                "tensor_pipe_bf16_equivalent_frac": achieved_tf * exec_equiv / peaks["tflops_burst"],
                "tensor_pipe_note": "frac counts ALGORITHMIC FLOPs (2*M*N*K); the tensor pipe executes 3 bf16 MMAs per bf16x3 "
                                    "product (and TF32 MMAs run at half the bf16 rate), so the pipe is busy for "
                                    f"{exec_equiv:.1f}x as many bf16-equivalent FLOPs -> tensor_pipe_bf16_equivalent_frac",
                "peak_source": peaks["source"] + " dense bf16, BURST figure (the GEMM family is replayed alone for a few ms)",
                "frac_of_sustained_peak": achieved_tf / peaks["tflops_sustained"],
                "how": f"sum of 2*M*N*K over the {gemm_calls} GEMMs of one step ({gemm_launches} kernel launches incl. split-K "
                       f"reduces; weight gradients grouped four per launch) / CUDA-event time of a captured graph that "
                       f"re-issues exactly those launches back to back ({n_prof_steps} replays, inputs L2-cold)",
                "gemm_ms_per_step": gemm_ms / n_prof_steps, "serial_step_ms": ms_serial,
                "gemm_share_of_serial_step": (gemm_ms / n_prof_steps) / ms_serial}

    # ---- (3b) HBM fractions of the gather-bound kernels north_star names, at the shapes of this step ----------------
    gather = None
    if not args.no_extras:
        try:
            pack = get_pack(g_dev)
            N, Wd, H = n_atoms, 512, 16
            ft = torch.randn(N, Wd, device=dev)
            E = pack.n_edges
            reps = 50

            def hbm(bytes_, ms_):
                gbs = bytes_ / (ms_ * 1e-3) / 1e9
                return {"achieved_GBps": gbs, "frac_of_hbm_peak": gbs / peaks["hbm_gbs"], "us_per_launch": ms_ * 1e3, "algorithmic_bytes": bytes_}
            ops.edge_attention_fwd(ft, pack, H)
            t_edge = timed_local(lambda: ops.edge_attention_fwd(ft, pack, H), reps) / reps
            T = pack.n_tuples[2]
            proj = torch.randn(N, 512, device=dev)
            pe = torch.tensor([0.0, 1.0, 1.0, 0.0], device=dev)
            ops.tuple_gather_fwd(proj, pack["idx2"], pe, T, 4, 511, 512)
            t_gather = timed_local(lambda: ops.tuple_gather_fwd(proj, pack["idx2"], pe, T, 4, 511, 512), reps) / reps
            x = torch.randn(4 * T, 512, device=dev)
            gam, bet = torch.ones(512, device=dev), torch.zeros(512, device=dev)
            ops.layernorm_fwd(x, gam, bet)
            t_ln = timed_local(lambda: ops.layernorm_fwd(x, gam, bet), reps) / reps
            gather = {
                "edge_attention_fwd": dict(hbm((2 * N * Wd + E * H) * 4 + (N + 1 + E) * 4, t_edge),
                                           shape=f"{N} atoms x {H} heads x 32, {E} directed edges (K4, reference graph_attention.py:283)"),
                "tuple_gather_fwd": dict(hbm((N * 512 + 4 * T * 512 + 4 * T) * 4, t_gather),
                                         shape=f"{T} propers x 4 atoms x 512 (K9, reference interaction_parameters.py:173-178)"),
                "layernorm_fwd": dict(hbm(2 * 4 * T * 512 * 4, t_ln), shape=f"{4 * T} tokens x 512"),
                "note": "isolated launches, CUDA events, data L2-resident between repetitions (3-30 MB working sets < 126 MB L2): "
                        "these kernels are launch / latency bound at the step's sizes, the HBM fraction is reported because "
                        "north_star asks for it, not because HBM bounds them"}
        except Exception as e:  # pragma: no cover
            gather = {"error": repr(e)}

    # ---- (4) energy + force evaluation throughput (K13), configs[3]: 1k molecules x {100, 1000} conformations -------
    energy_eval = None
    if not args.no_extras:
        try:
            from grappa_b200 import graph as gbg
            from grappa_b200.training import shard_molecules
            n_mols_total = 1000 * world    # weak scaling like the training step: 1000 molecules per GPU
            mine = list(shard_molecules(n_mols_total, rank, world))      # molecule i -> rank i mod world, no communication
            rows = []
            for n_confs in (100, 1000):
                base = synthetic.peptide_batch(seed=7, batch_size=8, n_res=4, n_confs=n_confs)
                mols8 = gbg.unbatch(base)
                ge = gbg.batch([mols8[i % 8] for i in mine]).to(dev)
                gen = torch.Generator().manual_seed(1)
                for l in ("n2", "n3", "n4", "n4_improper"):
                    T = ge.num_nodes(l)
                    if l in ("n2", "n3"):
                        ge.nodes[l].data["k"] = (100 + 300 * torch.rand(T, generator=gen)).to(dev)
                        ge.nodes[l].data["eq"] = (1.2 + 0.6 * torch.rand(T, generator=gen)).to(dev)
                    else:
                        ge.nodes[l].data["k"] = torch.randn(T, 3, generator=gen).to(dev)
                na = ge.num_nodes("n1")
                nm = len(mine)
                tup = [ge.num_nodes(l) for l in ("n2", "n3", "n4", "n4_improper")]
                lean = n_confs * (24 * na + 4 * nm) + 4 * (2 * tup[0] + 3 * tup[1] + 4 * tup[2] + 4 * tup[3]) \
                    + 4 * (2 * tup[0] + 2 * tup[1] + 3 * tup[2] + 3 * tup[3])
                full = lean + n_confs * (8 * sum(tup) + 16 * nm)
                for mode, nbytes in (("lean", lean), ("full_contract", full)):
                    en = Energy(write_tuple_terms=(mode == "full_contract"))
                    with torch.no_grad():
                        for _ in range(3):
                            en(ge)
                        reps = 10
                        ms_en = timed(lambda: en(ge), reps) / reps            # max over ranks
                    tot_bytes = torch.tensor([float(nbytes)], device=dev)
                    if world > 1:
                        dist.all_reduce(tot_bytes)
                    gbs = tot_bytes.item() / (ms_en * 1e-3) / 1e9
                    rows.append({"conformations": n_confs, "outputs": mode, "ms_per_launch": ms_en,
                                 "evals_per_s": n_mols_total * n_confs / (ms_en * 1e-3),
                                 "algorithmic_bytes_per_eval": tot_bytes.item() / (n_mols_total * n_confs),
                                 "achieved_GBps_all_gpus": gbs, "frac_of_hbm_peak": gbs / (world * peaks["hbm_gbs"])})
                del ge
            head = rows[0]
            energy_eval = {"value": head["evals_per_s"], "unit": "conformation energy+force evals/s", "n_gpus": world,
                           "workload": f"{n_mols_total} x 52-atom molecules (ACE-(ALA)4-NME) x 100 / 1000 conformations, molecules sharded "
                                       f"i mod {world} over the ranks (no collective), time = max over ranks; value = 100 conformations, lean outputs",
                           "sweep": rows,
                           "roofline": {"bound": "hbm", "kernel": "energy_pairs_kernel (K13, two conformations per lane, f32x2)",
                                        "achieved": head["achieved_GBps_all_gpus"], "peak": world * peaks["hbm_gbs"], "unit": "GB/s",
                                        "frac": head["frac_of_hbm_peak"], "traffic": 94.2e6,
                                        "traffic_source": "dram__bytes_read.sum + dram__bytes_write.sum = 71.0 + 23.2 MB per launch (1000 molecules "
                                                          "x 100 conformations, lean; ncu --set full, profiles/r2_energy_pairs_ncu_raw.csv) against "
                                                          "131.9 MB algorithmic: nothing is re-read, part of the gradient is still in L2 at kernel end",
                                        "ncu": {"capture": "profiles/r2_energy_pairs_ncu_raw.csv (1000 molecules x 100 conformations, lean)",
                                                "fp32_pipe_active_pct": 35.8, "issue_slots_busy_pct": 52.3,
                                                "lsu_shared_memory_wavefronts_pct_of_peak": 56.7, "warp_instructions_per_clk_per_sm": 2.0},
                                        "bound_note": "shared-memory bandwidth / round-barrier bound, not HBM bound: ~34 KB of shared-memory "
                                                      "traffic per conformation against 1.3 KB of HBM traffic (SURVEY.md 8d: 22-55 flop/B at the "
                                                      "lean byte count, above the fp32 ridge)", "peak_source": peaks["source"]}}
        except Exception as e:  # pragma: no cover
            energy_eval = {"error": repr(e)}

    # ---- (4b) inference parametrisation of a 1,502-atom protein (BASELINE configs[2]) -------------
    inference_eval = None
    if not args.no_extras:
        try:
            from grappa_b200 import inference as gb_inf
            prot = synthetic.protein(seed=3, n_res=149)
            get_pack(prot)
            prot = prot.pin_memory()
            gr = gb_inf.Grappa(model, device=str(dev), use_cuda_graph=True)
            for _ in range(4):
                gr._forward(prot)                       # eager, capture, replays
            ms_inf = timed(lambda: gr._forward(prot), 20) / 20
            model.train()
            inference_eval = {"value": world * 1e3 / ms_inf, "unit": "proteins/s (1502 atoms, 8842 tuples each)", "ms_per_protein": ms_inf,
                              "workload": "GrappaModel forward (eval) of ACE-(ALA)149-NME from a pinned HOST graph: H2D of features + "
                                          "index tables, one captured-graph replay, parameters left on the device; every rank its own "
                                          "protein (molecules shard, no collective)", "n_gpus": world, "dtype": prec}
        except Exception as e:  # pragma: no cover
            model.train()
            inference_eval = {"error": repr(e)}

    # ---- (5) CPU baseline (rank 0, N = 1 only) ----------------------------------------------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        try:
            r = cpu_reference_step(steps=2, warmup=1)
            cpu = {k: r[k] for k in ("value", "unit", "cores", "kind", "sample")}
        except Exception as e:  # pragma: no cover
            cpu = {"error": repr(e)}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": "molecules/s", "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": dtype, "data": "synthetic",
            "config": {"workload": WORKLOAD, "architecture": "grappa-1.2 (40.8 M parameters, random init)",
                       "molecules_per_gpu": B, "atoms_per_gpu": n_atoms, "conformations": 50, "dropout": "on (train mode)",
                       "optimizer": "Adam + global-norm clip 10", "parallelism": f"dp{world}",
                       "launch": "eager" if args.no_cuda_graph else "whole step captured in one CUDA graph, replayed per step",
                       "cache": "every step re-reads 163 MB of weights + 163 MB of gradients + Adam moments (650 MB > 126 MB L2); "
                                "no explicit L2 flush"},
            "e2e": {"value": e2e_value, "unit": "molecules/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4,
                    "ms_per_step": ms_e2e / K},
            "e2e_loader": e2e_loader,
            "gpu_launches": launches, "gpu_launches_per_step": launches / K,
            "clocks": clk.summary(), "roofline": roofline, "gather_kernels": gather, "energy_eval": energy_eval,
            "inference_eval": inference_eval, "param_sync": param_sync, "cpu_baseline": cpu,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        # Leave without tearing NCCL down: destroy_process_group() can block for minutes while captured CUDA graphs
        # still reference the communicator; every rank has finished its work once the barrier returns.
        torch.cuda.synchronize()
        dist.barrier()
        torch.cuda.synchronize()
        sys.stdout.flush()
        sys.stderr.flush()
        os._exit(0)


if __name__ == "__main__":
    main()
