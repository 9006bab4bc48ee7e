"""Developer probe: two batches in flight (two plans on two streams, alternating) against one stream."""
import copy, sys, time, torch
sys.path.insert(0, '/root/repo')
import hvit_b200
from hvit_b200.models import HybridViT
from hvit_b200.inference import AudioEnhancer

B, n = 64, 64000
m1 = HybridViT(precision="fp16").cuda().eval()
m2 = copy.deepcopy(m1)
e = [AudioEnhancer(m1, device="cuda"), AudioEnhancer(m2, device="cuda")]
x = [torch.randn(B, n, device="cuda") * 0.1 for _ in range(2)]
y = [torch.empty_like(x[0]) for _ in range(2)]
s = [torch.cuda.Stream(), torch.cuda.Stream()]
def run(K, two):
    torch.cuda.synchronize()
    t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
    t0.record()
    for st in s: st.wait_stream(torch.cuda.current_stream())
    for i in range(K):
        j = i % 2 if two else 0
        with torch.cuda.stream(s[j]):
            e[j].enhance_device(x[j], out=y[j])
    for st in s: torch.cuda.current_stream().wait_stream(st)
    t1.record(); torch.cuda.synchronize()
    return t0.elapsed_time(t1) / K
for two in (False, True): run(6, two)
for rep in range(3):
    for two in (False, True):
        ms = run(40, two)
        print(f"two_streams={two}: {ms:.3f} ms/step  {B * 4.0 / ms * 1e3:.0f} audio-s/s", flush=True)
