"""Generate the golden fixtures under tests/golden/ by running the REFERENCE itself.

Run in the build container only (it needs /root/reference, which does not exist
on the GPU box):

    python tests/golden/make_golden.py          # golden_v1 (tiny / 1 s / 2 s cases, literal-init forward)
    python tests/golden/make_golden.py --v2     # golden_v2 (4 s / 10 s, seeds 0-2, SNR 0/5/10 dB, W % 4 == 0, literal init)

What runs here is the reference's own code, unmodified:
  * ``models.HybridViT`` imported from /root/reference (models/hybrid_vit.py),
  * ``inference/enhancer.py`` executed verbatim; its two third-party imports
    that are not installed here (``librosa``, ``soundfile``) are satisfied by a
    shim whose ``stft`` / ``istft`` are backed by ``torch.stft`` / ``torch.istft``
    with librosa>=0.10 parameters (periodic Hann, center, zero padding) - an
    implementation independent of oracle/hvit_oracle.py's numpy/scipy one.

Weights and clips come from oracle.hvit_oracle.make_state_dict / synth_clip
(seeded, construction-order independent), so only outputs are stored.
"""
import importlib.util
import json
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"
sys.path.insert(0, ROOT)
sys.path.insert(0, REF)

from oracle import hvit_oracle as O  # noqa: E402

torch.set_num_threads(8)


def _install_shims():
    lib = types.ModuleType("librosa")

    def _stft(y, n_fft=2048, hop_length=None, win_length=None, window="hann", center=True, **kw):
        assert window == "hann"
        win_length = win_length or n_fft
        w = torch.hann_window(win_length, periodic=True, dtype=torch.float64)
        s = torch.stft(torch.from_numpy(np.asarray(y)).double(), n_fft, hop_length, win_length, w,
                       center=center, pad_mode="constant", return_complex=True)
        return s.numpy().astype(np.complex64 if np.asarray(y).dtype == np.float32 else np.complex128)

    def _istft(spec, hop_length=None, win_length=None, window="hann", center=True, length=None, **kw):
        n_fft = 2 * (spec.shape[0] - 1)
        win_length = win_length or n_fft
        w = torch.hann_window(win_length, periodic=True, dtype=torch.float64)
        y = torch.istft(torch.from_numpy(spec).to(torch.complex128), n_fft, hop_length, win_length, w,
                        center=center, length=length)
        return y.numpy().astype(np.float32 if spec.dtype == np.complex64 else np.float64)

    lib.stft, lib.istft = _stft, _istft
    sys.modules["librosa"] = lib
    sys.modules["soundfile"] = types.ModuleType("soundfile")


def _load_by_path(name, path):
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


TINY = dict(encoder_channels=[64, 64, 128], embed_dim=128, num_heads=2, num_layers=2,
            decoder_channels=[128, 64, 64, 1])
CASES = {
    # name: (cfg overrides, weight seed, clip seconds, clip seed, n_samples override)
    "tiny_0p5s": (TINY, 11, 0.5, 3, None),
    "tiny_ragged": (TINY, 12, None, 4, 9001),      # n not a multiple of hop
    "default_1s": ({}, 0, 1.0, 0, None),
    "default_2s": ({}, 1, 2.0, 1, None),
}
# golden_v2: the headline clip length (4 s), the longest latency-sweep length (10 s, N = 1248 tokens), three weight
# seeds, SNR 0 / 5 / 10 dB, a clip whose last encoder map is a multiple of the patch size (T = 512 -> W = 128 = 4 * 32;
# the other cases leave a remainder of 1..3 columns), and the reference's LITERAL initialisation (BatchNorm 0/1
# statistics, saturated tanh) through the whole enhance path.
#   name: (cfg overrides, weight seed | "literal", clip seconds, clip seed, n_samples override, snr_db)
CASES_V2 = {
    "default_4s_s0_snr0": ({}, 0, 4.0, 10, None, 0.0),
    "default_4s_s1_snr10": ({}, 1, 4.0, 11, None, 10.0),
    "default_4s_s2_snr5": ({}, 2, 4.0, 12, None, 5.0),
    "default_10s": ({}, 0, 10.0, 13, None, 5.0),
    "default_w128": ({}, 1, None, 14, 511 * 128, 5.0),
    "literal_1s": ({}, "literal", 1.0, 15, None, 5.0),
}
STAGE_SAMPLES = 1024
MODEL_OUT_SAMPLES = 65536   # v2 stores the model output at seeded random positions (+ its max |.|), not the full map


def run_cases(HybridViT, enh_mod, cases, v2):
    out = {}
    meta = {}
    for name, spec in cases.items():
        over, wseed, secs, cseed, nsamp = spec[:5]
        snr_db = spec[5] if v2 else 5.0
        cfg = O.full_cfg(over)
        kwargs = {k: cfg[k] for k in ("encoder_channels", "embed_dim", "num_heads", "num_layers", "decoder_channels")}
        if wseed == "literal":      # the reference's own _init_weights under torch.manual_seed(0)
            torch.manual_seed(0)
            model = HybridViT(**kwargs).eval()
            sd = {k: v.detach().clone() for k, v in model.state_dict().items()}
        else:
            sd = O.make_state_dict(cfg, seed=wseed)
            model = HybridViT(**kwargs).eval()
            model.load_state_dict(sd, strict=True)
        clean, noisy = O.synth_clip(seconds=secs or 1.0, seed=cseed, n_samples=nsamp, snr_db=snr_db)
        enhancer = enh_mod.AudioEnhancer(model, device="cpu")
        y = enhancer.enhance(noisy, normalize=True)
        # forward-only golden on the normalised magnitude the enhancer feeds the model
        dbg = {}
        O.enhance(sd, noisy, cfg, debug=dbg)
        x = torch.from_numpy(dbg["noisy_mag_norm"]).float()[None, None]
        stages = {}
        hooks = []
        names = {f"encoder.{i}": model.encoder[i] for i in range(3)}
        names["transformer"] = model.transformer
        names["to_feature_map"] = model.to_feature_map
        for i in range(4):
            names[f"decoder.{i}"] = model.decoder[i]
        for l in range(cfg["num_layers"]):
            names[f"transformer.blocks.{l}"] = model.transformer.blocks[l]
        for n, m in names.items():
            hooks.append(m.register_forward_hook(lambda mod, inp, o, n=n: stages.__setitem__(n, o.detach())))
        with torch.no_grad():
            fwd = model(x)
        for h in hooks:
            h.remove()
        with torch.no_grad():
            fwd2, attns = model(x, return_attentions=True)
        assert torch.equal(fwd, fwd2)
        out[f"{name}/waveform"] = np.asarray(y, dtype=np.float32)
        if v2:
            flat = fwd.reshape(-1).numpy()
            idx = np.random.default_rng(4321).integers(0, flat.size, size=MODEL_OUT_SAMPLES)
            # (positions are default_rng(4321).integers(0, size, MODEL_OUT_SAMPLES): regenerated by the tests)
            out[f"{name}/model_out_val"] = flat[idx]
            out[f"{name}/model_out_absmax"] = np.float32(np.abs(flat).max())
        else:
            out[f"{name}/model_out"] = fwd.squeeze().numpy()
        out[f"{name}/attn0_head0_row0"] = attns[0][0, 0, 0].numpy()
        rng = np.random.default_rng(1234)
        for sn, t in stages.items():
            # to_feature_map hook fires on [B,N,C]; store as the oracle does ([B,C,H,W]) later in the test
            flat = t.reshape(-1).numpy()
            idx = rng.integers(0, flat.size, size=min(STAGE_SAMPLES, flat.size))
            out[f"{name}/stage/{sn}/idx"] = idx.astype(np.int64)
            out[f"{name}/stage/{sn}/val"] = flat[idx]
        meta[name] = dict(cfg=over, weight_seed=wseed, seconds=secs, clip_seed=cseed, n_samples=nsamp,
                          weights_sha256=O.state_dict_digest(sd), T=int(x.shape[-1]),
                          out_std=float(fwd.std()), sisdr_clean_vs_ref=O.si_sdr(clean, y))
        if v2:
            meta[name]["snr_db"] = snr_db
        print(name, meta[name])
    return out, meta


def main():
    _install_shims()
    from models import HybridViT  # reference
    enh_mod = _load_by_path("ref_enhancer", os.path.join(REF, "inference", "enhancer.py"))
    if "--v2" in sys.argv:
        out, meta = run_cases(HybridViT, enh_mod, CASES_V2, v2=True)
        np.savez_compressed(os.path.join(HERE, "golden_v2.npz"), **out)
        with open(os.path.join(HERE, "golden_v2.json"), "w") as f:
            json.dump(meta, f, indent=1, sort_keys=True)
        print("wrote", os.path.getsize(os.path.join(HERE, "golden_v2.npz")), "bytes")
        return
    out, meta = run_cases(HybridViT, enh_mod, CASES, v2=False)
    # literal reference init: digest only (28M params are not stored)
    torch.manual_seed(0)
    ref0 = HybridViT().eval()
    meta["literal_init_seed0"] = dict(weights_sha256=O.state_dict_digest(ref0.state_dict()))
    xin = torch.rand(2, 1, 257, 63, generator=torch.Generator().manual_seed(5))
    with torch.no_grad():
        out["literal_init_seed0/model_out"] = ref0(xin).numpy()
    np.savez_compressed(os.path.join(HERE, "golden_v1.npz"), **out)
    with open(os.path.join(HERE, "golden_v1.json"), "w") as f:
        json.dump(meta, f, indent=1, sort_keys=True)
    print("wrote", os.path.getsize(os.path.join(HERE, "golden_v1.npz")), "bytes")


if __name__ == "__main__":
    main()
