"""The packed, device-resident weight set the CUDA plan consumes (include/hvit.h: hvit_weights), derived from a
reference-keyed state_dict by ``hvit_pack_weights`` (csrc/pack.cu, C ABI, no torch arithmetic): BatchNorm folding,
K-major conv re-layout, 16-bit conversion and the pre-summed 2x2 parity kernels for "nearest x2 upsample + 3x3 conv"
decoder blocks.  ``PackedWeights`` only stages the fp32 state_dict on the device, allocates the packed buffer and keeps
it alive.  Runs once per weight version.

``fold_bn`` / ``conv_khwc`` / ``up2_parity_kernels`` are the same re-expressions written with torch ops; the tests use
them (in fp64 on the CPU, and against the CUDA packer's output on the GPU) - the product path does not call them.
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
    """Owns the packed device buffer and the ctypes view (``.c``) handed to hvit_plan_create."""

    def __init__(self, model, precision: int, c_cfg=None):
        import ctypes as C
        lib = _lib.load()
        dev = next(model.parameters()).device
        cfg = model.arch
        ccfg = c_cfg if c_cfg is not None else model._c_cfg(precision)
        sd = {k: v.detach().to(device=dev, dtype=torch.float32).contiguous()
              for k, v in model.state_dict().items() if v.is_floating_point()}
        ref = _lib.RefWeights()
        for i in range(len(cfg["encoder_channels"])):
            ref.enc_conv_w[i] = sd[f"encoder.{i}.block.0.weight"].data_ptr()
            bn = f"encoder.{i}.block.1"
            ref.enc_bn_w[i], ref.enc_bn_b[i] = sd[f"{bn}.weight"].data_ptr(), sd[f"{bn}.bias"].data_ptr()
            ref.enc_bn_mean[i], ref.enc_bn_var[i] = sd[f"{bn}.running_mean"].data_ptr(), sd[f"{bn}.running_var"].data_ptr()
        ref.patch_w = sd["patch_embed.projection.weight"].data_ptr()
        ref.patch_b = sd["patch_embed.projection.bias"].data_ptr()
        pos = sd["pos_encoding.pos_embed"]
        ref.pos_embed, ref.pos_len = pos.data_ptr(), int(pos.shape[1])
        for l in range(cfg["num_layers"]):
            p = f"transformer.blocks.{l}"
            ref.ln1_w[l], ref.ln1_b[l] = sd[f"{p}.norm1.weight"].data_ptr(), sd[f"{p}.norm1.bias"].data_ptr()
            ref.ln2_w[l], ref.ln2_b[l] = sd[f"{p}.norm2.weight"].data_ptr(), sd[f"{p}.norm2.bias"].data_ptr()
            ref.qkv_w[l], ref.qkv_b[l] = sd[f"{p}.attn.qkv.weight"].data_ptr(), sd[f"{p}.attn.qkv.bias"].data_ptr()
            ref.proj_w[l], ref.proj_b[l] = sd[f"{p}.attn.proj.weight"].data_ptr(), sd[f"{p}.attn.proj.bias"].data_ptr()
            ref.fc1_w[l], ref.fc1_b[l] = sd[f"{p}.mlp.net.0.weight"].data_ptr(), sd[f"{p}.mlp.net.0.bias"].data_ptr()
            ref.fc2_w[l], ref.fc2_b[l] = sd[f"{p}.mlp.net.3.weight"].data_ptr(), sd[f"{p}.mlp.net.3.bias"].data_ptr()
        ref.lnf_w, ref.lnf_b = sd["transformer.norm.weight"].data_ptr(), sd["transformer.norm.bias"].data_ptr()
        ref.tofm_w, ref.tofm_b = sd["to_feature_map.weight"].data_ptr(), sd["to_feature_map.bias"].data_ptr()
        n_dec = len(cfg["decoder_channels"])
        for i in range(n_dec):
            ci = 1 if cfg["decoder_upsample_factors"][i] > 1 else 0
            ref.dec_conv_w[i] = sd[f"decoder.{i}.block.{ci}.weight"].data_ptr()
            if i == n_dec - 1:
                continue
            bn = f"decoder.{i}.block.{ci + 1}"
            ref.dec_bn_w[i], ref.dec_bn_b[i] = sd[f"{bn}.weight"].data_ptr(), sd[f"{bn}.bias"].data_ptr()
            ref.dec_bn_mean[i], ref.dec_bn_var[i] = sd[f"{bn}.running_mean"].data_ptr(), sd[f"{bn}.running_var"].data_ptr()
            if cfg["use_skip_connections"] and f"skip_projections.{i}.weight" in sd:
                ref.skip_w[i] = sd[f"skip_projections.{i}.weight"].data_ptr()
                ref.skip_b[i] = sd[f"skip_projections.{i}.bias"].data_ptr()
        nbytes = lib.hvit_packed_weights_bytes(C.byref(ccfg), ref.pos_len)
        if nbytes == 0:
            _lib.check(-1, "hvit_packed_weights_bytes")
        with torch.cuda.device(dev):
            raw = torch.empty(nbytes + 256, dtype=torch.uint8, device=dev)
            shift = (-raw.data_ptr()) % 256
            self.buffer = raw[shift:shift + nbytes]
            self.c = _lib.Weights()
            _lib.check(lib.hvit_pack_weights(C.byref(ccfg), C.byref(ref), self.buffer.data_ptr(), nbytes, C.byref(self.c),
                                             _lib.current_stream_ptr()), "hvit_pack_weights")
            # the fp32 staging copies (`sd`) are released when this returns: wait for the packing kernels that read them
            torch.cuda.current_stream().synchronize()
        self.nbytes = nbytes
