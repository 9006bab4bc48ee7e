"""CPU: pin the oracle (oracle/hvit_oracle.py) against outputs of the reference
itself (tests/golden/, produced by tests/golden/make_golden.py) and against
independent STFT implementations."""
import numpy as np
import pytest
import torch

CASES = ["tiny_0p5s", "tiny_ragged", "default_1s", "default_2s"]


def _case(oracle, meta, name):
    m = meta[name]
    cfg = oracle.full_cfg(m["cfg"])
    sd = oracle.make_state_dict(cfg, seed=m["weight_seed"])
    clean, noisy = oracle.synth_clip(seconds=m["seconds"] or 1.0, seed=m["clip_seed"], n_samples=m["n_samples"])
    return cfg, sd, clean, noisy


@pytest.mark.parametrize("name", CASES)
def test_weights_reproducible(oracle, golden, name):
    _, meta = golden
    cfg, sd, _, _ = _case(oracle, meta, name)
    assert oracle.state_dict_digest(sd) == meta[name]["weights_sha256"]


@pytest.mark.parametrize("name", CASES)
def test_oracle_matches_reference_enhance(oracle, golden, name):
    arrays, meta = golden
    cfg, sd, clean, noisy = _case(oracle, meta, name)
    dbg = {}
    y = oracle.enhance(sd, noisy, cfg, debug=dbg)
    ref_y = arrays[f"{name}/waveform"]
    assert y.shape == ref_y.shape and y.dtype == np.float32
    # fp32 tolerance of the north star for the fp32 mode
    assert oracle.max_rel_err(y, ref_y) <= 1e-4
    assert oracle.max_rel_err(dbg["model_out"], arrays[f"{name}/model_out"]) <= 1e-4
    assert abs(oracle.si_sdr(clean, y) - meta[name]["sisdr_clean_vs_ref"]) <= 0.05


@pytest.mark.parametrize("name", CASES)
def test_oracle_stages_match_reference(oracle, golden, name):
    arrays, meta = golden
    cfg, sd, _, noisy = _case(oracle, meta, name)
    dbg = {}
    oracle.enhance(sd, noisy, cfg, debug=dbg)
    x = torch.from_numpy(dbg["noisy_mag_norm"]).float()[None, None]
    stages = {}
    with torch.no_grad():
        out, attns = oracle.hybrid_vit_forward(sd, x, cfg, return_attentions=True, stages=stages)
    B, C, H, W = stages["to_feature_map"].shape
    # the reference hook sees to_feature_map as [B, N, C]
    stages["to_feature_map"] = stages["to_feature_map"].reshape(B, C, H * W).transpose(1, 2)
    checked = 0
    for key in arrays.files:
        if key.startswith(f"{name}/stage/") and key.endswith("/idx"):
            sn = key.split("/")[2]
            idx = arrays[key]
            val = arrays[key[:-3] + "val"]
            ours = stages[sn].reshape(-1).numpy()[idx]
            scale = max(np.abs(val).max(), 1e-6)
            assert np.abs(ours - val).max() / scale <= 2e-5, sn
            checked += 1
    assert checked >= 9
    np.testing.assert_allclose(attns[0][0, 0, 0].numpy(), arrays[f"{name}/attn0_head0_row0"], rtol=1e-4, atol=1e-7)


CASES_V2 = ["default_4s_s0_snr0", "default_4s_s1_snr10", "default_4s_s2_snr5", "default_10s", "default_w128", "literal_1s"]


@pytest.mark.parametrize("name", CASES_V2)
def test_oracle_matches_reference_v2(oracle, golden2, name):
    """golden_v2: headline 4 s clips (three weight seeds, SNR 0/5/10 dB), 10 s (N = 1248 tokens), a W % 4 == 0 length
    and the reference's literal initialisation - waveform, model output (seeded sample), SI-SDR and per-stage samples."""
    from conftest import golden_case, golden_model_out_err
    arrays, meta = golden2
    m = meta[name]
    cfg, sd, clean, noisy = golden_case(oracle, m)
    assert oracle.state_dict_digest(sd) == m["weights_sha256"]
    dbg = {}
    y = oracle.enhance(sd, noisy, cfg, debug=dbg)
    ref_y = arrays[f"{name}/waveform"]
    assert y.shape == ref_y.shape and dbg["model_out"].shape[1] == m["T"]
    assert oracle.max_rel_err(y, ref_y) <= 1e-4
    assert golden_model_out_err(oracle, arrays, name, dbg["model_out"]) <= 1e-4
    assert abs(oracle.si_sdr(clean, y) - m["sisdr_clean_vs_ref"]) <= 0.05
    if name in ("default_4s_s0_snr0", "default_w128"):   # per-stage samples (one remainder-1 and one W % 4 == 0 case)
        x = torch.from_numpy(dbg["noisy_mag_norm"]).float()[None, None]
        stages = {}
        with torch.no_grad():
            oracle.hybrid_vit_forward(sd, x, cfg, stages=stages)
        B, C, H, W = stages["to_feature_map"].shape
        stages["to_feature_map"] = stages["to_feature_map"].reshape(B, C, H * W).transpose(1, 2)
        checked = 0
        for key in arrays.files:
            if key.startswith(f"{name}/stage/") and key.endswith("/idx"):
                sn = key.split("/")[2]
                val = arrays[key[:-3] + "val"]
                ours = stages[sn].reshape(-1).numpy()[arrays[key]]
                assert np.abs(ours - val).max() / max(np.abs(val).max(), 1e-6) <= 2e-5, sn
                checked += 1
        assert checked >= 9


def test_stft_against_torch(oracle):
    _, noisy = oracle.synth_clip(seconds=1.3, seed=7, n_samples=20777)
    s = oracle.stft(noisy)
    assert s.shape == (257, 1 + 20777 // 128) and s.dtype == np.complex64
    w = torch.hann_window(512, periodic=True, dtype=torch.float64)
    t = torch.stft(torch.from_numpy(noisy).double(), 512, 128, 512, w, center=True, pad_mode="constant",
                   return_complex=True).numpy()
    assert np.abs(s - t).max() <= 2e-6 * np.abs(t).max()
    y = oracle.istft(s, length=len(noisy))
    ty = torch.istft(torch.from_numpy(s).to(torch.complex128), 512, 128, 512, w, center=True,
                     length=len(noisy)).numpy()
    assert y.dtype == np.float32 and y.shape == noisy.shape
    assert np.abs(y - ty).max() <= 1e-6
    # round trip; the ragged tail (n % hop samples, covered only by the centre padding) is exact too
    assert np.abs(y - noisy).max() <= 1e-5


def test_istft_envelope_is_1p5_inside(oracle):
    wss = oracle.window_sumsquare(20, 512, 128, 512)
    assert np.allclose(wss[512:-512], 1.5, atol=1e-6)


def test_enhance_guards(oracle):
    cfg = oracle.full_cfg(dict(encoder_channels=[64, 64, 128], embed_dim=128, num_heads=2, num_layers=1,
                               decoder_channels=[128, 64, 64, 1]))
    sd = oracle.make_state_dict(cfg, seed=3)
    y = oracle.enhance(sd, np.zeros(4096, dtype=np.float32), cfg)   # both <=1e-8 guards fire
    assert y.shape == (4096,) and np.isfinite(y).all()


def test_literal_init_forward(oracle, golden):
    """Oracle forward on the reference's own literal init (seed 0) matches the
    reference output; uses an nn.Module-free path (weights regenerated by the
    product's HybridViT mirror in test_boundary.py)."""
    arrays, meta = golden
    assert "literal_init_seed0" in meta and arrays["literal_init_seed0/model_out"].shape == (2, 1, 257, 63)


def test_metrics_match_reference_golden():
    """hvit_b200.evaluation.metrics against values produced by the reference's own evaluation/metrics.py
    (tests/golden/make_golden_metrics.py, same seeded signals)."""
    import json
    import os
    import hvit_b200  # noqa: F401
    from hvit_b200.evaluation import metrics as M
    with open(os.path.join(os.path.dirname(__file__), "golden", "metrics_v1.json")) as f:
        cases = json.load(f)["cases"]
    rng = np.random.default_rng(123)
    for case in cases:
        n = case["n"]
        t = np.arange(n) / 16000.0
        clean = (0.3 * np.sin(2 * np.pi * 220 * t) * (0.5 + 0.5 * np.sin(2 * np.pi * 2 * t))
                 + 0.01 * rng.standard_normal(n)).astype(np.float32)
        noisy = (clean + 0.1 * rng.standard_normal(n)).astype(np.float32)
        enh = (0.9 * clean + 0.02 * rng.standard_normal(n)).astype(np.float32)
        assert abs(M.compute_sisdr(clean, enh) - case["sisdr"]) < 1e-4
        assert abs(M.compute_snr(clean, enh) - case["snr"]) < 1e-4
        assert abs(M.compute_segsnr(clean, enh) - case["segsnr"]) < 1e-4
        assert abs(M.compute_lsd(clean, enh) - case["lsd"]) < 2e-3      # (torch.stft shim vs numpy rfft, fp32 logs)
        assert abs(M.compute_sisdr(clean, noisy) - case["sisdr_noisy"]) < 1e-4
        assert abs(M.compute_snr(clean, noisy) - case["snr_noisy"]) < 1e-4
        allm = M.compute_all_metrics(clean, enh, noisy)
        assert set(allm) == {"pesq", "stoi", "sisdr", "snr", "segsnr", "lsd", "pesq_improvement", "stoi_improvement",
                             "sisdr_improvement", "snr_improvement"}


def test_loss_goldens_are_reproduced_by_the_formulas_the_device_path_uses():
    """tests/golden/losses_v1.json (values from the reference's training/losses.py) against the closed forms the CUDA
    reduction implements (L1 / MSE means, 1 - cosine similarity), in fp64 on the CPU: pins the arithmetic of the
    validation forward without a GPU."""
    import json
    import os
    import torch
    gold = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "losses_v1.json")))

    def pair(seed, shape):
        g = torch.Generator().manual_seed(seed)
        target = torch.rand(*shape, generator=g)
        pred = (target + 0.1 * torch.randn(*shape, generator=g)).clamp_min(0.0)
        return pred.double(), target.double()

    for c in gold["cases"]:
        if c["kind"] != "combined":
            continue
        kw = dict(l1_weight=1.0, mse_weight=0.0, stoi_weight=0.1, perceptual_weight=0.0, use_log_compression=False)
        kw.update(c["kwargs"])
        p, t = pair(c["seed"], c["shape"])
        a, b = (torch.log(p.float() + 1e-8).double(), torch.log(t.float() + 1e-8).double()) if kw["use_log_compression"] else (p, t)
        cos = (p.flatten(1) * t.flatten(1)).sum(1) / (p.flatten(1).norm(dim=1).clamp_min(1e-12) * t.flatten(1).norm(dim=1).clamp_min(1e-12))
        total = (kw["l1_weight"] * (a - b).abs().mean() + kw["mse_weight"] * ((a - b) ** 2).mean()
                 + kw["stoi_weight"] * (1.0 - cos).mean() + kw["perceptual_weight"] * (p - t).abs().mean())
        assert abs(float(total) - c["total"]) <= 2e-6 * max(1.0, abs(c["total"])), (c["kwargs"], float(total), c["total"])
