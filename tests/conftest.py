import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def golden():
    d = os.path.join(ROOT, "tests", "golden")
    arrays = np.load(os.path.join(d, "golden_v1.npz"))
    with open(os.path.join(d, "golden_v1.json")) as f:
        meta = json.load(f)
    return arrays, meta


@pytest.fixture(scope="session")
def golden2():
    """golden_v2: 4 s / 10 s clips, weight seeds 0-2, SNR 0/5/10 dB, W % 4 == 0, literal init (make_golden.py --v2)."""
    d = os.path.join(ROOT, "tests", "golden")
    arrays = np.load(os.path.join(d, "golden_v2.npz"))
    with open(os.path.join(d, "golden_v2.json")) as f:
        meta = json.load(f)
    return arrays, meta


def golden_case(oracle, m):
    """(cfg, state_dict, clean, noisy) of a golden case from its metadata; "literal" weights are the reference's own
    initialisation under torch.manual_seed(0), reproduced by the module mirror (digest-checked by the callers)."""
    import torch
    cfg = oracle.full_cfg(m["cfg"])
    if m["weight_seed"] == "literal":
        from hvit_b200.models import HybridViT
        torch.manual_seed(0)
        kw = {k: cfg[k] for k in ("encoder_channels", "embed_dim", "num_heads", "num_layers", "decoder_channels")}
        sd = {k: v.detach().clone() for k, v in HybridViT(**kw).state_dict().items()}
    else:
        sd = oracle.make_state_dict(cfg, seed=m["weight_seed"])
    clean, noisy = oracle.synth_clip(seconds=m["seconds"] or 1.0, seed=m["clip_seed"], n_samples=m["n_samples"],
                                     snr_db=m.get("snr_db", 5.0))
    return cfg, sd, clean, noisy


def golden_model_out_err(oracle, arrays, name, ours_model_out):
    """max-rel error of a full model output against the seeded sample of the reference's (golden_v2)."""
    flat = np.asarray(ours_model_out).reshape(-1)
    idx = np.random.default_rng(4321).integers(0, flat.size, size=arrays[f"{name}/model_out_val"].size)
    d = np.abs(flat[idx].astype(np.float64) - arrays[f"{name}/model_out_val"].astype(np.float64)).max()
    return float(d / float(arrays[f"{name}/model_out_absmax"]))


@pytest.fixture(scope="session")
def oracle():
    from oracle import hvit_oracle
    return hvit_oracle
