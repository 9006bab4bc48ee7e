"""HybridViT with the reference's constructor, state_dict keys and forward signature
(reference models/hybrid_vit.py:21-525), executed by the sm_100a CUDA plan in csrc/.

The module tree below only holds parameters (fp32 masters, so ``state_dict()`` /
``load_state_dict(strict=True)`` round-trip with reference checkpoints).  ``forward``
packs the weights once per parameter version, builds one launch plan per input shape and
enqueues it on the current CUDA stream; no PyTorch op touches the activations.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Dict, List, Optional

import torch
import torch.nn as nn

from .. import _lib
from .attention import VisionTransformer
from .components import ConvBlock, PatchEmbedding, PositionalEncoding, TransposeConvBlock
from .packing import PackedWeights

_PRECISIONS = {"fp16": _lib.PREC_FP16, "bf16": _lib.PREC_BF16, "fp32": _lib.PREC_FP32}


class _Plan:
    """One hvit_plan (+ its workspace) for a fixed (B, F, T, n_samples, precision)."""

    def __init__(self, lib, cfg: _lib.ModelCfg, weights: PackedWeights, B: int, F: int, T: int, n_samples: int,
                 device: torch.device, debug: bool = False):
        self.lib = lib
        nbytes = lib.hvit_workspace_bytes(C.byref(cfg), B, F, T, n_samples)
        if nbytes == 0:
            _lib.check(-1, "hvit_workspace_bytes")
        raw = torch.empty(nbytes + 1024, dtype=torch.uint8, device=device)
        shift = (-raw.data_ptr()) % 1024
        self.workspace = raw[shift:shift + nbytes]
        self.weights = weights  # keeps the packed tensors alive as long as the plan
        handle = C.c_void_p()
        # setup kernels go on torch's current stream: ordered after the weight packing, and before this stream's forwards
        _lib.check(lib.hvit_plan_create(C.byref(cfg), C.byref(weights.c), B, F, T, n_samples,
                                        self.workspace.data_ptr(), nbytes, _lib.current_stream_ptr(), C.byref(handle)),
                   "hvit_plan_create")
        self.handle = handle
        self.debug = bool(debug)
        if self.debug:
            _lib.check(lib.hvit_plan_set_debug(handle, 1), "hvit_plan_set_debug")
        self.B, self.F, self.T, self.n_samples = B, F, T, n_samples
        self.act_dtype = torch.float16 if cfg.precision == _lib.PREC_FP16 else torch.bfloat16
        hp, wp = C.c_int(), C.c_int()
        self.tokens = lib.hvit_plan_tokens(handle, C.byref(hp), C.byref(wp))
        self.grid = (hp.value, wp.value)

    def buffer(self, name: str) -> torch.Tensor:
        """Typed view of an internal activation buffer (tests / debugging)."""
        off, dims, es = C.c_size_t(), (C.c_int * 4)(), C.c_int()
        rank = self.lib.hvit_plan_buffer(self.handle, name.encode(), C.byref(off), C.byref(dims), C.byref(es))
        if rank < 0:
            _lib.check(rank, f"hvit_plan_buffer({name})")
        shape = [dims[i] for i in range(rank)]
        dt = {2: self.act_dtype, 4: torch.float32, 8: torch.complex64}[es.value]
        if name in ("max_val", "mag_max"):
            dt = torch.float32
        n = 1
        for s in shape:
            n *= s
        return self.workspace[off.value:off.value + n * es.value].view(dt).view(shape)

    def launch_count(self, enhance: bool) -> int:
        return self.lib.hvit_plan_launch_count(self.handle, 1 if enhance else 0)

    def steps(self, enhance: bool = True):
        """[{name, kernel, algo_flops, exec_flops, algo_bytes, launches}] for measurement (bench.py)."""
        e = 1 if enhance else 0
        out = []
        for i in range(self.lib.hvit_plan_num_steps(self.handle, e)):
            name, kern = C.create_string_buffer(64), C.create_string_buffer(32)
            af, ef, ab, nl = C.c_double(), C.c_double(), C.c_double(), C.c_int()
            _lib.check(self.lib.hvit_plan_step_info(self.handle, e, i, name, 64, kern, 32, C.byref(af), C.byref(ef),
                                                    C.byref(ab), C.byref(nl)), "hvit_plan_step_info")
            out.append(dict(name=name.value.decode(), kernel=kern.value.decode(), algo_flops=af.value,
                            exec_flops=ef.value, algo_bytes=ab.value, launches=nl.value))
        return out

    def enhance_profiled(self, wave_in: torch.Tensor, wave_out: torch.Tensor, normalize: bool = True):
        """One enhance call with a CUDA event after every step; returns per-step milliseconds (synchronises)."""
        n = self.lib.hvit_plan_num_steps(self.handle, 1)
        ms = (C.c_float * n)()
        _lib.check(self.lib.hvit_enhance_profiled(self.handle, wave_in.data_ptr(), wave_out.data_ptr(),
                                                  1 if normalize else 0, _lib.current_stream_ptr(), ms, n),
                   "hvit_enhance_profiled")
        return list(ms)

    def __del__(self):
        try:
            if self.handle:
                self.lib.hvit_plan_destroy(self.handle)
                self.handle = None
        except Exception:
            pass


class HybridViT(nn.Module):
    """CNN encoder -> patch embedding -> ViT -> CNN decoder with U-Net skips (reference hybrid_vit.py:21-170).

    Extra keyword (not in the reference): ``precision`` =
      ``"fp16"`` (default) tcgen05 tensor cores, fp16 operands/activations, fp32 accumulate - the 16-bit parity mode
                           (meets the north star's 1e-2 / 0.05 dB bar);
      ``"fp32"``           CUDA-core accuracy mode (<= 1e-4);
      ``"bf16"``           same kernels as fp16 with bf16 operands, kept for dynamic range only - NOT a parity mode
                           (0.6-2.2e-2 max-rel, the bf16 rounding floor).
    An explicit ``precision=`` wins over the ``HVIT_PRECISION`` environment default.
    Supported architecture subset, checked at plan creation:
    1 input / 1 output channel, 3x3 convolutions, pool sizes and upsample factors in {1, 2},
    head_dim 64, channel counts multiples of 64 (bf16) / 16 (fp32), no CLS token.
    """

    def __init__(self, input_channels: int = 1, output_channels: int = 1,
                 encoder_channels: List[int] = [64, 128, 256], encoder_kernel_sizes: List[int] = [3, 3, 3],
                 encoder_pool_sizes: List[int] = [2, 2, 1], embed_dim: int = 512, num_heads: int = 8,
                 num_layers: int = 6, mlp_ratio: float = 4.0, patch_size: int = 4,
                 decoder_channels: List[int] = [256, 128, 64, 1], decoder_kernel_sizes: List[int] = [3, 3, 3, 3],
                 decoder_upsample_factors: List[int] = [1, 2, 2, 1], dropout: float = 0.1, attn_dropout: float = 0.1,
                 drop_path_rate: float = 0.1, use_skip_connections: bool = True, use_cls_token: bool = False,
                 precision: Optional[str] = None):
        super().__init__()
        if use_cls_token:
            raise NotImplementedError("use_cls_token=True is not wired to any reference config and is not supported")
        if input_channels != 1 or output_channels != 1:
            raise NotImplementedError("hvit_b200 supports the magnitude-spectrogram model: 1 input / 1 output channel")
        if any(k != 3 for k in list(encoder_kernel_sizes) + list(decoder_kernel_sizes)):
            raise NotImplementedError("only 3x3 convolutions are supported")
        self.input_channels, self.output_channels = input_channels, output_channels
        self.embed_dim, self.patch_size = embed_dim, patch_size
        self.use_skip_connections, self.use_cls_token = use_skip_connections, use_cls_token
        self.arch = dict(encoder_channels=list(encoder_channels), encoder_pool_sizes=list(encoder_pool_sizes),
                         embed_dim=embed_dim, num_heads=num_heads, num_layers=num_layers, mlp_ratio=mlp_ratio,
                         patch_size=patch_size, decoder_channels=list(decoder_channels),
                         decoder_upsample_factors=list(decoder_upsample_factors),
                         use_skip_connections=use_skip_connections)
        # an explicit precision= wins; HVIT_PRECISION only supplies the default
        self.precision = precision if precision is not None else os.environ.get("HVIT_PRECISION", "fp16")
        self.debug_buffers = False   # True: plans also store the test-only intermediates ("model_out", "logits")

        # --- same registration order as the reference so the literal init consumes the RNG identically
        self.encoder = nn.ModuleList()
        cin = input_channels
        for ch, ks, pool in zip(encoder_channels, encoder_kernel_sizes, encoder_pool_sizes):
            self.encoder.append(ConvBlock(cin, ch, kernel_size=ks, padding=ks // 2,
                                          pool_size=pool if pool > 1 else None, activation="relu",
                                          use_batchnorm=True, dropout=dropout))
            cin = ch
        self.patch_embed = PatchEmbedding(cin, embed_dim, patch_size=patch_size, flatten=True)
        self.cls_token = None
        self.pos_encoding = PositionalEncoding(embed_dim, max_len=10000, learnable=True, dropout=dropout)
        self.transformer = VisionTransformer(embed_dim, num_layers=num_layers, num_heads=num_heads, mlp_ratio=mlp_ratio,
                                             qkv_bias=True, dropout=dropout, attn_dropout=attn_dropout,
                                             drop_path_rate=drop_path_rate)
        self.to_feature_map = nn.Linear(embed_dim, cin)
        self.decoder = nn.ModuleList()
        n_dec = len(decoder_channels)
        for i, (ch, ks, up) in enumerate(zip(decoder_channels, decoder_kernel_sizes, decoder_upsample_factors)):
            ic = decoder_channels[0] if i == 0 else decoder_channels[i - 1]
            last = i == n_dec - 1
            if use_skip_connections and not last:
                ic += ch
            self.decoder.append(TransposeConvBlock(ic, ch, kernel_size=ks, padding=ks // 2,
                                                   upsample_factor=up if up > 1 else None, activation="relu",
                                                   use_batchnorm=True, dropout=0.0 if last else dropout,
                                                   final_layer=last))
        if use_skip_connections:
            self.skip_projections = nn.ModuleList([
                nn.Conv2d(ec, dc, kernel_size=1) for ec, dc in zip(encoder_channels[::-1], decoder_channels[:-1])])
        else:
            self.skip_projections = None
        self.apply(self._init_weights)

        self._packed: Dict[int, tuple] = {}
        self._plans: Dict[tuple, _Plan] = {}   # LRU: key -> plan, most recently used last
        self.max_plans = 16

    # reference hybrid_vit.py:265-284
    @staticmethod
    def _init_weights(m: nn.Module) -> None:
        if isinstance(m, nn.Linear):
            nn.init.trunc_normal_(m.weight, std=0.02)
            if m.bias is not None:
                nn.init.constant_(m.bias, 0)
        elif isinstance(m, nn.Conv2d):
            nn.init.kaiming_normal_(m.weight, mode="fan_out", nonlinearity="relu")
            if m.bias is not None:
                nn.init.constant_(m.bias, 0)
        elif isinstance(m, (nn.BatchNorm2d, nn.LayerNorm)):
            nn.init.constant_(m.weight, 1)
            nn.init.constant_(m.bias, 0)

    # ------------------------------------------------------------------ plan plumbing
    def _c_cfg(self, precision: int) -> _lib.ModelCfg:
        a = self.arch
        c = _lib.ModelCfg()
        c.n_enc = len(a["encoder_channels"])
        for i, (ch, pool) in enumerate(zip(a["encoder_channels"], a["encoder_pool_sizes"])):
            c.enc_channels[i], c.enc_pool[i] = ch, max(int(pool), 1)
        c.embed_dim, c.num_heads, c.num_layers = a["embed_dim"], a["num_heads"], a["num_layers"]
        c.mlp_hidden, c.patch_size = int(a["embed_dim"] * a["mlp_ratio"]), a["patch_size"]
        c.n_dec = len(a["decoder_channels"])
        for i, (ch, up) in enumerate(zip(a["decoder_channels"], a["decoder_upsample_factors"])):
            c.dec_channels[i], c.dec_up[i] = ch, max(int(up), 1)
        c.use_skip = 1 if a["use_skip_connections"] else 0
        c.precision = precision
        c.ln_eps = 1e-5
        return c

    def _weights_version(self) -> tuple:
        # (in-place version counter, storage address) of every parameter / buffer.  The tensor list is cached - walking
        # the module tree on every call costs more than a whole single-clip enhance - and refreshed by _apply()
        # (.to / .cuda / .half ...) and load_state_dict(); replace a submodule's Parameter object by hand and you must
        # call model._refresh_version_tensors() yourself.
        ts = getattr(self, "_version_tensors", None)
        if ts is None:
            ts = self._refresh_version_tensors()
        return tuple(t._version for t in ts) + tuple(t.data_ptr() for t in ts)

    def _refresh_version_tensors(self):
        ts = list(self.state_dict(keep_vars=True).values())
        object.__setattr__(self, "_version_tensors", ts)
        return ts

    def _apply(self, fn, *args, **kwargs):
        out = super()._apply(fn, *args, **kwargs)
        self._refresh_version_tensors()
        return out

    def load_state_dict(self, *args, **kwargs):
        out = super().load_state_dict(*args, **kwargs)
        self._refresh_version_tensors()
        return out

    def _get_packed(self, precision: int) -> PackedWeights:
        ver = self._weights_version()
        hit = self._packed.get(precision)
        if hit is None or hit[0] != ver:
            self._packed[precision] = (ver, PackedWeights(self, precision))
            self._plans = {k: v for k, v in self._plans.items() if k[0] != precision}
        return self._packed[precision][1]

    def plan_for(self, B: int, F: int, T: int, n_samples: int = 0) -> _Plan:
        """Launch plan for a batch shape (built on first use, cached)."""
        if self.training:
            raise RuntimeError("hvit_b200.HybridViT is inference-only: call model.eval() first")
        if self.precision not in _PRECISIONS:
            raise ValueError(f"precision must be one of {sorted(_PRECISIONS)}, got {self.precision!r}")
        dev = next(self.parameters()).device
        if dev.type != "cuda":
            raise RuntimeError("hvit_b200.HybridViT runs on a CUDA (sm_100) device only; move the model with .to('cuda')")
        lib = _lib.load()
        prec = _PRECISIONS[self.precision]
        with torch.cuda.device(dev):
            packed = self._get_packed(prec)
            key = (prec, B, F, T, n_samples, bool(self.debug_buffers))
            plan = self._plans.pop(key, None)
            if plan is None:
                plan = _Plan(lib, self._c_cfg(prec), packed, B, F, T, n_samples, dev, debug=self.debug_buffers)
                while len(self._plans) >= self.max_plans:     # least recently used first (dicts keep insertion order)
                    self._plans.pop(next(iter(self._plans)))
            self._plans[key] = plan                           # (re-)insert as most recently used
        return plan

    # ------------------------------------------------------------------ forward (reference hybrid_vit.py:396-469)
    @torch.no_grad()
    def forward(self, x: torch.Tensor, return_attentions: bool = False):
        if x.dim() != 4 or x.shape[1] != 1:
            raise ValueError(f"expected input [B, 1, F, T], got {tuple(x.shape)}")
        if not x.is_cuda:
            raise RuntimeError("hvit_b200.HybridViT.forward needs a CUDA tensor (no CPU path)")
        x = x.contiguous().float()
        B, _, F, T = x.shape
        plan = self.plan_for(B, F, T)
        lib = plan.lib
        y = torch.empty_like(x)
        probs = None
        if return_attentions:
            L, h, N = self.arch["num_layers"], self.arch["num_heads"], plan.tokens
            probs = torch.empty((L, B, h, N, N), dtype=torch.float32, device=x.device)
        with torch.cuda.device(x.device):
            _lib.check(lib.hvit_forward(plan.handle, x.data_ptr(), y.data_ptr(), _lib.ptr(probs),
                                        _lib.current_stream_ptr()), "hvit_forward")
        if return_attentions:
            return y, [probs[l] for l in range(probs.shape[0])]
        return y

    def count_parameters(self) -> Dict[str, int]:
        """reference hybrid_vit.py:471-489"""
        n = lambda mod: sum(p.numel() for p in mod.parameters())  # noqa: E731
        return {"encoder": n(self.encoder), "transformer": n(self.transformer), "decoder": n(self.decoder),
                "total": n(self), "trainable": sum(p.numel() for p in self.parameters() if p.requires_grad)}


def create_hybrid_vit(config: Optional[Dict] = None) -> HybridViT:
    """Build from the merged YAML dict; same keys, defaults and quirks as the reference
    (hybrid_vit.py:492-525): every dropout comes from ``encoder.dropout``, ``use_cls_token``
    is not configurable, ``strides`` keys are ignored.  ``model.precision`` (optional) is ours."""
    m = (config or {}).get("model", {})
    enc, tr, dec = m.get("encoder", {}), m.get("transformer", {}), m.get("decoder", {})
    return HybridViT(
        input_channels=m.get("input_channels", 1), output_channels=m.get("output_channels", 1),
        encoder_channels=enc.get("channels", [64, 128, 256]), encoder_kernel_sizes=enc.get("kernel_sizes", [3, 3, 3]),
        encoder_pool_sizes=enc.get("pool_sizes", [2, 2, 1]),
        embed_dim=tr.get("embed_dim", 512), num_heads=tr.get("num_heads", 8), num_layers=tr.get("num_layers", 6),
        mlp_ratio=tr.get("mlp_ratio", 4), patch_size=tr.get("patch_size", 4),
        decoder_channels=dec.get("channels", [256, 128, 64, 1]), decoder_kernel_sizes=dec.get("kernel_sizes", [3, 3, 3, 3]),
        decoder_upsample_factors=dec.get("upsample_factors", [1, 2, 2, 1]),
        dropout=enc.get("dropout", 0.1), attn_dropout=tr.get("attention_dropout", 0.1),
        drop_path_rate=tr.get("drop_path_rate", 0.1), use_skip_connections=dec.get("use_skip_connections", True),
        precision=m.get("precision"))
