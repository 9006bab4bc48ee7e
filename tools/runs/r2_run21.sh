#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x -k "return_attentions or losses or validator or validate_loop or folded or layernorm" > gpurun_out/r2t_tests.log 2>&1
echo "rc=$?" >> gpurun_out/r2t_tests.log
tail -25 gpurun_out/r2t_tests.log
python - <<'P'
import sys, time, torch
sys.path.insert(0, '/root/repo')
import hvit_b200
from hvit_b200.models import HybridViT
m = HybridViT(precision="fp16").cuda().eval()
x = torch.rand(16, 1, 257, 501, device='cuda')
for ra in (False, True):
    for _ in range(2): m(x, return_attentions=ra)
    torch.cuda.synchronize(); t = time.perf_counter()
    for _ in range(5): m(x, return_attentions=ra)
    torch.cuda.synchronize(); print("return_attentions", ra, "bs16 x 4 s forward ms:", (time.perf_counter() - t) / 5 * 1e3)
P
