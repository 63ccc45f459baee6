"""Kernel timeline of one captured training step (CUPTI through torch.profiler) -> gpurun_out/step_trace.json
plus a text summary: wall time, per-stream busy time, idle gaps, top kernels by summed duration."""
import json, os, sys, collections
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import profile, ProfilerActivity
from grappa_b200 import models, ops, synthetic
from grappa_b200.energy import Energy
from grappa_b200.loss import MolwiseLoss
from grappa_b200.pack import get_pack
from grappa_b200.training import Trainer

eager = "--eager" in sys.argv
if "--serial" in sys.argv:      # one stream: every kernel alone on the GPU -> clean per-kernel durations
    from grappa_b200 import tape as _tape
    _tape.set_concurrency(False)
ops.set_matmul_precision(os.environ.get("GRAPPA_B200_PREC", ops.BENCH_PRECISION))
from grappa_b200.training import init_distributed
rank, local, world = init_distributed()
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
torch.manual_seed(0)
model = models.model_from_config(models.grappa_1_2_model_config()).train()
tr = Trainer(model, Energy(write_tuple_terms=False), MolwiseLoss(gradient_weight=0.8, energy_weight=1.0, param_weight=0.0,
             proper_regularisation=1e-3, improper_regularisation=1e-3), lr=1.5e-5, clip=10.0, device=dev, use_cuda_graph=not eager)
g = synthetic.peptide_batch(seed=100, batch_size=32, n_res=4, n_confs=50)
get_pack(g)
g = g.pin_memory()
for _ in range(5):
    tr.step(g)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for _ in range(3):
        tr.step(g)
    torch.cuda.synchronize()
if rank != 0:
    torch.cuda.synchronize()
    import torch.distributed as dist
    dist.barrier()
    os._exit(0)
out = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out", os.environ.get("GRAPPA_B200_TRACE", "step_trace.json"))
prof.export_chrome_trace(out)
ev = json.load(open(out))["traceEvents"]
ks = [e for e in ev if e.get("cat") == "kernel"]
ks.sort(key=lambda e: e["ts"])
print("kernels recorded:", len(ks))
# split into steps by the adam kernel
ends = [i for i, e in enumerate(ks) if "adam" in e["name"]]
lo = ends[0] + 1 if len(ends) >= 2 else 0
hi = ends[1] + 1 if len(ends) >= 2 else len(ks)
step = ks[lo:hi]
t0 = min(e["ts"] for e in step); t1 = max(e["ts"] + e["dur"] for e in step)
print(f"step: {len(step)} kernels, wall {t1 - t0:.0f} us, sum of kernel durations {sum(e['dur'] for e in step):.0f} us")
# union busy time
iv = sorted((e["ts"], e["ts"] + e["dur"]) for e in step)
busy = 0.0; cs, ce = iv[0]
for s, e in iv[1:]:
    if s > ce:
        busy += ce - cs; cs, ce = s, e
    else:
        ce = max(ce, e)
busy += ce - cs
print(f"GPU busy (union of kernels) {busy:.0f} us, idle {t1 - t0 - busy:.0f} us")
by_stream = collections.defaultdict(float)
for e in step:
    by_stream[e["args"].get("stream")] += e["dur"]
print("busy per stream:", {k: round(v) for k, v in sorted(by_stream.items(), key=lambda kv: -kv[1])})
agg = collections.defaultdict(lambda: [0, 0.0])
for e in step:
    n = e["name"].split("(")[0].replace("void ", "")[:60]
    agg[n][0] += 1; agg[n][1] += e["dur"]
for n, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:45]:
    print(f"{t:9.1f} us {c:5d}  {t / c:7.1f} us/launch  {n}")
# phases on the main timeline: coarse 0.5 ms buckets of concurrency (sum of durations / bucket width)
B = 500.0
nb = int((t1 - t0) / B) + 1
occ = [0.0] * nb
for e in step:
    s, d = e["ts"] - t0, e["dur"]
    while d > 0:
        b = int(s / B); take = min(d, (b + 1) * B - s)
        occ[b] += take; s += take; d -= take
print("avg concurrent kernels per 0.5 ms bucket:", [round(o / B, 2) for o in occ])
nccl = [e for e in step if "nccl" in e["name"].lower()]
if nccl:
    print("NCCL kernels of the step (start us, duration us):", [(round(e["ts"] - t0), round(e["dur"])) for e in nccl])
tail = [e for e in step if any(k in e["name"] for k in ("sumsq", "adam", "tick"))]
print("optimizer tail (name, start us, duration us):", [(e["name"].split("(")[0][-24:], round(e["ts"] - t0), round(e["dur"])) for e in tail])
other = [e for e in step if e not in tail and "nccl" not in e["name"].lower()]
print("last compute kernel before the optimizer ends at", round(max(e["ts"] + e["dur"] for e in other) - t0), "us:",
      max(other, key=lambda e: e["ts"] + e["dur"])["name"][:60])
if world > 1:
    import torch.distributed as dist
    dist.barrier()
    os._exit(0)
