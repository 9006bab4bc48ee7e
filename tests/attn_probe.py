"""Developer probe: time the tcgen05 attention kernel on the model's shape."""
import sys, torch
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
import gpu_util as U
from hvit_b200 import _lib
for (B, N, h) in [(64, 496, 8), (64, 1248, 8), (8, 112, 8), (512, 128, 8)]:
    qkv = torch.randn(B * N, 3 * h * 64, device='cuda').half()
    out = torch.empty(B * N, h * 64, device='cuda', dtype=torch.float16)
    f = lambda: _lib.check(U.lib().hvit_attention_16(U.P(qkv), U.P(out), B, N, h, 1, U.stream()), "attn")
    for _ in range(3): f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20): f()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 20
    q, k, v = qkv.float().view(B, N, 3, h, 64).permute(2, 0, 3, 1, 4)
    ref = torch.nn.functional.scaled_dot_product_attention(q, k, v).permute(0, 2, 1, 3).reshape(B * N, h * 64)
    err = (out.float() - ref).abs().max().item() / ref.abs().max().item()
    print(f"B={B} N={N} h={h}: {ms*1e3:.1f} us  {4*B*h*N*N*64/ms/1e9:.1f} TF/s  max-rel err {err:.2e}", flush=True)
