"""Developer probe: folded-LayerNorm consumer GEMMs against the plain ones on the model's shapes."""
import sys, torch
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
import gpu_util as U
from hvit_b200 import _lib
M, D = 31744, 512
x = torch.randn(M, D, device='cuda')
x16 = torch.empty(M, D, device='cuda', dtype=torch.float16)
S = D // 128
stats = torch.empty(M, S, 2, device='cuda')
_lib.check(U.lib().hvit_rowstats_16(U.P(x), U.P(x16), U.P(stats), M, D, S, 1, U.stream()), "rowstats")
g = torch.ones(D, device='cuda'); be = torch.zeros(D, device='cuda')
def timeit(f, reps=20):
    for _ in range(3): f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): f()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3
for name, N, act in (("qkv", 1536, 0), ("fc1", 2048, 2)):
    w = (torch.randn(N, D, device='cuda') * 0.05).half(); b = torch.randn(N, device='cuda')
    out = torch.empty(M, N, device='cuda', dtype=torch.float16)
    wsc = torch.empty(N, D, device='cuda', dtype=torch.float16); gc = torch.empty(2 * N, device='cuda')
    plain = timeit(lambda: U.gemm_16(x16, w, shift=b, act=act, out=out))
    fold = timeit(lambda: _lib.check(U.lib().hvit_linear_ln_consumer_16(U.P(x16), U.P(stats), S, U.P(w), U.P(g), U.P(be), U.P(b), 1e-5, act,
                                     U.P(out), N, M, N, D, 1, U.P(wsc), U.P(gc), U.stream()), "c"))
    print(f"{name}: plain {plain:.1f} us, folded consumer (incl. the weight-fold kernel) {fold:.1f} us", flush=True)
