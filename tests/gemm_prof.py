"""Developer probe: per-role cycle counters of the CTA-pair GEMM (HVIT_PROF=1), one launch per transformer shape."""
import os, sys, torch
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
os.environ["HVIT_PROF"] = "1"
import gpu_util as U
M = 31744
for name, (N, K, f32, res, act) in dict(qkv=(1536, 512, False, False, 0), proj=(512, 512, True, True, 0),
                                         fc1=(2048, 512, False, False, 2), fc2=(512, 2048, True, True, 0),
                                         big=(4096, 4096, False, False, 0)).items():
    a = torch.randn(M, K, device='cuda').half(); w = (torch.randn(N, K, device='cuda') * 0.05).half()
    bias = torch.randn(N, device='cuda')
    out = torch.empty(M, N, device='cuda', dtype=torch.float32 if f32 else torch.float16)
    for _ in range(2):
        U.gemm_16(a, w, shift=bias, act=act, residual=out if res else None, out_f32=f32, out=out)
    torch.cuda.synchronize()
