"""Dev: GPU busy time vs wall time of the fused step (torch.profiler / kineto kernel timeline)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "gs-lora_b200"), ROOT, os.path.join(ROOT, "tests")]
import torch
from torch.profiler import profile, ProfilerActivity
sys.argv = ["bench"]
import bench, engine_cl
dev = torch.device("cuda", 0)
model, cfg = bench.build_model(dev)
g = torch.Generator().manual_seed(1)
t = [torch.rand(512, 3, 112, 112, generator=g).to(dev), torch.randint(0, 100, (512,), generator=g).to(dev),
     torch.rand(512, 3, 112, 112, generator=g).to(dev), torch.randint(0, 100, (512,), generator=g).to(dev)]
kw = dict(beta=0.15, alpha=1e-4, BND=105.0, hparams=dict(lr=1e-2, wd=0.05))
for _ in range(3):
    engine_cl.unlearn_step(model, *t, **kw)
torch.cuda.synchronize()
steps = 5
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        engine_cl.unlearn_step(model, *t, **kw)
    e1.record(); torch.cuda.synchronize()
wall = e0.elapsed_time(e1) / steps
evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
busy = sum(e.device_time for e in evs) / 1e3 / steps if evs and hasattr(evs[0], "device_time") else sum(e.cuda_time for e in evs) / 1e3 / steps
print(f"wall {wall:.2f} ms/step, GPU busy (sum of kernel+memcpy durations) {busy:.2f} ms/step, idle {wall - busy:.2f} ms ({100 * (wall - busy) / wall:.1f}%)")
agg = {}
for e in evs:
    d = e.device_time if hasattr(e, "device_time") else e.cuda_time
    agg.setdefault(e.name[:60], [0, 0.0]); agg[e.name[:60]][0] += 1; agg[e.name[:60]][1] += d
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:16]:
    print(f"  {v[1] / 1e3 / steps:7.3f} ms  n={v[0] / steps:5.1f}  {k}")

# sync vs pipelined read-back, same process, alternating blocks of 20 steps
def run(fn, n=20):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    prev = None
    for _ in range(n):
        r = fn()
        if prev is not None and hasattr(prev, "wait"):
            prev.wait()
        prev = r
    if hasattr(prev, "wait"):
        prev.wait()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
for rep in range(3):
    a = run(lambda: engine_cl.unlearn_step(model, *t, **kw))
    b = run(lambda: engine_cl.unlearn_step_async(model, *t, **kw))
    print(f"sync {a:.2f} ms/step   pipelined read-back {b:.2f} ms/step")
