"""Developer probe: LayerNorm kernel rate with an L2-resident vs an HBM-resident input."""
import sys, torch
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
import gpu_util as U
from hvit_b200 import _lib
D = 512
g = torch.ones(D, device='cuda'); b = torch.zeros(D, device='cuda')
flush = torch.empty(512 << 20, dtype=torch.uint8, device='cuda')
for rows in (2048, 8192, 16384, 31744, 79872):
    x = torch.randn(rows, D, device='cuda'); out = torch.empty(rows, D, device='cuda', dtype=torch.float16)
    f = lambda: _lib.check(U.lib().hvit_layernorm(U.P(x), U.P(g), U.P(b), U.P(out), 2, rows, D, 1e-5, U.stream()), "ln")
    for _ in range(3): f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20): f()
    e1.record(); torch.cuda.synchronize()
    hot = e0.elapsed_time(e1) / 20
    cold = 0.0
    for _ in range(5):
        flush.zero_(); e0.record(); f(); e1.record(); torch.cuda.synchronize(); cold += e0.elapsed_time(e1) / 5
    mb = rows * D * 6 / 1e6
    print(f"rows={rows}: back-to-back {hot*1e3:.1f} us ({mb/hot/1e3:.2f} TB/s), after L2 flush {cold*1e3:.1f} us ({mb/cold/1e3:.2f} TB/s)", flush=True)
