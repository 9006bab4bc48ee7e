"""One tiny-model enhance + one forward (both 16-bit and fp32 kernels, varlen path included) for compute-sanitizer:
   compute-sanitizer --tool memcheck|racecheck|synccheck|initcheck python tools/sanitize_smoke.py"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import hvit_oracle as O          # noqa: E402  (seeded weights / clips only)
import hvit_b200                              # noqa: E402,F401
from hvit_b200.models import HybridViT       # noqa: E402
from hvit_b200.inference import AudioEnhancer  # noqa: E402

over = dict(encoder_channels=[64, 64, 128], embed_dim=128, num_heads=2, num_layers=2, decoder_channels=[128, 64, 64, 1])
cfg = O.full_cfg(over)
sd = O.make_state_dict(cfg, seed=7)
for precision in (sys.argv[1:] or ["fp16", "fp32"]):
    m = HybridViT(precision=precision, **{k: cfg[k] for k in ("encoder_channels", "embed_dim", "num_heads", "num_layers",
                                                                "decoder_channels")})
    m.load_state_dict(sd, strict=True)
    enh = AudioEnhancer(m.cuda().eval(), device="cuda:0")
    _, noisy = O.synth_clip(seconds=0.5, seed=7)
    y = enh.enhance(noisy)
    ref = O.enhance(sd, noisy, cfg)
    yv = enh.enhance_varlen([noisy, noisy[:5000], noisy[:2047]])
    x = torch.rand(2, 1, 257, 70).cuda()
    out, attn = m(x, return_attentions=True)
    torch.cuda.synchronize()
    print(precision, "enhance max-rel", O.max_rel_err(y, ref), "varlen[0]==single", bool(np.array_equal(yv[0], y)),
          "forward", tuple(out.shape), flush=True)
print("sanitize smoke done")
