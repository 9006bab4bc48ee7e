"""Derive the packed, device-resident weight set the CUDA plan consumes (include/hvit.h: hvit_weights)
from a reference-keyed state_dict: BatchNorm folding, NHWC / K-major re-layout, bf16 conversion, and the
pre-summed 2x2 parity kernels for "nearest x2 upsample + 3x3 conv" decoder blocks.

Runs once per weight version on the device with torch tensor ops (host-side plumbing, not the hot path).
"""
from __future__ import annotations

from typing import Dict, List

import torch

from .. import _lib

BN_EPS = 1e-5  # nn.BatchNorm2d default used by the reference (components.py:67,162)

# R[p][a][k]: which of the 3 kernel rows k collapse onto low-resolution tap a for output parity p
_PARITY = torch.tensor([[[1.0, 0.0, 0.0], [0.0, 1.0, 1.0]],
                        [[1.0, 1.0, 0.0], [0.0, 0.0, 1.0]]])


def fold_bn(sd: Dict[str, torch.Tensor], prefix: str):
    w, b = sd[f"{prefix}.weight"].float(), sd[f"{prefix}.bias"].float()
    mean, var = sd[f"{prefix}.running_mean"].float(), sd[f"{prefix}.running_var"].float()
    scale = w / torch.sqrt(var + BN_EPS)
    return scale.contiguous(), (b - mean * scale).contiguous()


def conv_khwc(w: torch.Tensor) -> torch.Tensor:
    """[Cout, Cin, kh, kw] -> [Cout, kh, kw, Cin] (K-major rows for the implicit GEMM)."""
    return w.permute(0, 2, 3, 1).contiguous()


def up2_parity_kernels(w: torch.Tensor) -> torch.Tensor:
    """[Cout, Cin, 3, 3] -> [4, Cout, 2, 2, Cin]: out[2y+py, 2x+px] = sum_{a,b} K[py*2+px][:, a, b, :] . in[y+a+py-1, x+b+px-1]."""
    R = _PARITY.to(w.device, torch.float32)
    k = torch.einsum("pak,oikl,qbl->pqoabi", R, w.float(), R)  # [2,2,Cout,2,2,Cin]
    return k.reshape(4, w.shape[0], 2, 2, w.shape[1]).contiguous()


class PackedWeights:
    """Owns the packed tensors and the ctypes view handed to hvit_plan_create."""

    def __init__(self, model, precision: int):
        dev = next(model.parameters()).device
        sd = {k: v.detach() for k, v in model.state_dict().items()}
        act = {_lib.PREC_BF16: torch.bfloat16, _lib.PREC_FP16: torch.float16, _lib.PREC_FP32: torch.float32}[precision]
        self.keep: List[torch.Tensor] = []
        self.c = _lib.Weights()
        c = self.c
        cfg = model.arch

        def hold(t: torch.Tensor, dtype=torch.float32) -> int:
            t = t.to(device=dev, dtype=dtype).contiguous()
            self.keep.append(t)
            return t.data_ptr()

        # encoder: block 0 is the stem (fp32 CUDA-core kernel), the rest are implicit GEMMs
        for i in range(len(cfg["encoder_channels"])):
            w = sd[f"encoder.{i}.block.0.weight"]
            scale, shift = fold_bn(sd, f"encoder.{i}.block.1")
            if i == 0:
                c.stem_w = hold(w[:, 0].permute(1, 2, 0))           # [3,3,C0]
                c.stem_scale, c.stem_shift = hold(scale), hold(shift)
            elif precision != _lib.PREC_FP32:
                # 16-bit modes: the BN scale is folded into the conv weights in fp32 (one rounding to 16 bits), the
                # epilogue only adds the shift (no per-channel scale loads; scale pointer = NULL means 1.0)
                c.enc_w[i] = hold(conv_khwc(w.float() * scale.to(w.device)[:, None, None, None]), act)
                c.enc_shift[i] = hold(shift)
            else:
                c.enc_w[i] = hold(conv_khwc(w), act)
                c.enc_scale[i], c.enc_shift[i] = hold(scale), hold(shift)
        c.patch_w = hold(conv_khwc(sd["patch_embed.projection.weight"]), act)
        c.patch_b = hold(sd["patch_embed.projection.bias"])
        pos = sd["pos_encoding.pos_embed"]
        c.pos_embed = hold(pos.reshape(pos.shape[1], pos.shape[2]))
        c.pos_len = int(pos.shape[1])
        for l in range(cfg["num_layers"]):
            p = f"transformer.blocks.{l}"
            c.ln1_g[l], c.ln1_b[l] = hold(sd[f"{p}.norm1.weight"]), hold(sd[f"{p}.norm1.bias"])
            c.ln2_g[l], c.ln2_b[l] = hold(sd[f"{p}.norm2.weight"]), hold(sd[f"{p}.norm2.bias"])
            c.qkv_w[l], c.qkv_b[l] = hold(sd[f"{p}.attn.qkv.weight"], act), hold(sd[f"{p}.attn.qkv.bias"])
            c.proj_w[l], c.proj_b[l] = hold(sd[f"{p}.attn.proj.weight"], act), hold(sd[f"{p}.attn.proj.bias"])
            c.fc1_w[l], c.fc1_b[l] = hold(sd[f"{p}.mlp.net.0.weight"], act), hold(sd[f"{p}.mlp.net.0.bias"])
            c.fc2_w[l], c.fc2_b[l] = hold(sd[f"{p}.mlp.net.3.weight"], act), hold(sd[f"{p}.mlp.net.3.bias"])
        c.lnf_g, c.lnf_b = hold(sd["transformer.norm.weight"]), hold(sd["transformer.norm.bias"])
        c.tofm_w, c.tofm_b = hold(sd["to_feature_map.weight"], act), hold(sd["to_feature_map.bias"])
        n_dec = len(cfg["decoder_channels"])
        for i in range(n_dec):
            up = cfg["decoder_upsample_factors"][i]
            ci = 1 if up > 1 else 0
            w = sd[f"decoder.{i}.block.{ci}.weight"]
            if i == n_dec - 1:
                c.head_w = hold(w[0].permute(1, 2, 0))              # [3,3,C]
                continue
            scale, shift = fold_bn(sd, f"decoder.{i}.block.{ci + 1}")
            if precision != _lib.PREC_FP32:  # BN scale folded into the weights (see the encoder above)
                ws = w.float() * scale.to(w.device)[:, None, None, None]
                c.dec_w[i] = hold(up2_parity_kernels(ws) if up > 1 else conv_khwc(ws), act)
                c.dec_shift[i] = hold(shift)
            else:
                c.dec_w[i] = hold(conv_khwc(w), act)
                c.dec_scale[i], c.dec_shift[i] = hold(scale), hold(shift)
            if cfg["use_skip_connections"] and f"skip_projections.{i}.weight" in sd:
                sw = sd[f"skip_projections.{i}.weight"]
                c.skip_w[i] = hold(sw.reshape(sw.shape[0], sw.shape[1]), act)
                c.skip_b[i] = hold(sd[f"skip_projections.{i}.bias"])
