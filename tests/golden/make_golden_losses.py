"""Generates tests/golden/losses_v1.json by running the REFERENCE's own training/losses.py (loaded by file path) and
the loss-averaging loop of training/trainer.py:207-251 on seeded spectrogram pairs.  Run in the build container only:
    python tests/golden/make_golden_losses.py
The inputs are regenerated from the seeds by the tests (torch.Generator, CPU)."""
import importlib.util
import json
import os

import torch

HERE = os.path.dirname(os.path.abspath(__file__))


def pair(seed, shape):
    g = torch.Generator().manual_seed(seed)
    target = torch.rand(*shape, generator=g)
    pred = (target + 0.1 * torch.randn(*shape, generator=g)).clamp_min(0.0)
    return pred, target


def main():
    spec = importlib.util.spec_from_file_location("ref_losses", "/root/reference/training/losses.py")
    ref = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref)
    cases = []
    configs = [dict(), dict(l1_weight=0.5, mse_weight=2.0, stoi_weight=0.0), dict(use_log_compression=True, mse_weight=1.0),
               dict(perceptual_weight=0.3, use_log_compression=True), dict(l1_weight=0.0, mse_weight=0.0, stoi_weight=1.0)]
    for ci, kw in enumerate(configs):
        for seed, shape in ((1, (2, 1, 257, 63)), (2, (3, 1, 257, 126)), (3, (1, 1, 64, 17))):
            pred, target = pair(seed, shape)
            total, comps = ref.CombinedLoss(**kw)(pred, target, return_components=True)
            cases.append(dict(kind="combined", kwargs=kw, seed=seed, shape=list(shape), total=float(total),
                              components={k: float(v) for k, v in comps.items()}))
    for lt in ("l1", "mse", "l1+mse"):
        for red in ("mean", "sum"):
            for log in (False, True):
                pred, target = pair(7, (2, 1, 257, 40))
                v = ref.SpectrogramLoss(loss_type=lt, reduction=red, use_log_compression=log)(pred, target)
                cases.append(dict(kind="spectrogram", loss_type=lt, reduction=red, use_log_compression=log, seed=7,
                                  shape=[2, 1, 257, 40], value=float(v)))
    for red in ("mean", "sum"):
        pred, target = pair(9, (4, 1, 257, 33))
        cases.append(dict(kind="stoi", reduction=red, seed=9, shape=[4, 1, 257, 33],
                          value=float(ref.STOILoss(reduction=red)(pred, target))))
    # the validation loop's averaging (trainer.py:229-249): mean over batches of criterion(pred, target)
    crit = ref.create_loss_function({"loss": {"l1_weight": 1.0, "mse_weight": 0.0, "stoi_weight": 0.1}})
    tot, nb = 0.0, 0
    for seed, shape in ((11, (2, 1, 257, 63)), (12, (2, 1, 257, 63)), (13, (1, 1, 257, 63))):
        pred, target = pair(seed, shape)
        tot += crit(pred, target).item()
        nb += 1
    cases.append(dict(kind="validate_loop", seeds=[11, 12, 13], shapes=[[2, 1, 257, 63], [2, 1, 257, 63], [1, 1, 257, 63]],
                      loss=tot / nb))
    with open(os.path.join(HERE, "losses_v1.json"), "w") as f:
        json.dump(dict(generator="tests/golden/make_golden_losses.py", reference="training/losses.py", cases=cases), f, indent=1)
    print(len(cases), "cases")


if __name__ == "__main__":
    main()
