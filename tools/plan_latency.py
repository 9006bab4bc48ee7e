"""Developer probe: what an unseen (batch, length) costs - plan creation (geometry, workspace allocation, TMA maps, one-time
weight-derived buffers) and the first call on it - against a steady-state call."""
import json, sys, time, torch
sys.path.insert(0, '/root/repo')
import numpy as np
import hvit_b200
from hvit_b200.models import HybridViT
from hvit_b200.inference import AudioEnhancer

m = HybridViT(precision="fp16").cuda().eval()
enh = AudioEnhancer(m, device="cuda")
enh.enhance(np.random.randn(16000).astype(np.float32))   # library / tables / weight packing warm
torch.cuda.synchronize()
rows = []
for B, sec in [(1, 2.0), (1, 3.5), (1, 7.0), (8, 2.5), (64, 1.5), (64, 3.0), (64, 6.0)]:
    n = int(sec * 16000)
    x = torch.randn(B, n, device="cuda")
    torch.cuda.synchronize(); t0 = time.perf_counter()
    plan = m.plan_for(B, 257, 1 + n // 128, n)
    torch.cuda.synchronize(); t1 = time.perf_counter()
    y = enh.enhance_device(x)
    torch.cuda.synchronize(); t2 = time.perf_counter()
    for _ in range(3): enh.enhance_device(x)
    torch.cuda.synchronize(); t3 = time.perf_counter()
    for _ in range(10): enh.enhance_device(x)
    torch.cuda.synchronize(); t4 = time.perf_counter()
    ws = None
    for attr in ("workspace", "ws", "_ws"):
        t = getattr(plan, attr, None)
        if isinstance(t, torch.Tensor):
            ws = int(t.numel() * t.element_size())
    rows.append(dict(batch=B, seconds=sec, plan_create_ms=(t1 - t0) * 1e3, first_call_ms=(t2 - t1) * 1e3,
                     steady_ms=(t4 - t3) / 10 * 1e3, workspace_mb=None if ws is None else ws / 2**20))
    print(rows[-1], flush=True)
json.dump(rows, open("gpurun_out/r2_plan_latency.json", "w"), indent=1)
