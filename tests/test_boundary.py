"""CPU (-m "not gpu"): the drop-in boundary without a GPU - the C-ABI library loads and exports every declared
symbol, the module mirror has the reference's state_dict keys / literal init, configs load, host-side plan
geometry and error behaviour, and the data-parallel sharding logic (gloo, world_size 2)."""
import ctypes as C
import os
import re
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    import __graft_entry__ as g
    g.build()  # compiles for sm_100a when stale (nvcc cross-compiles without a GPU)
    from hvit_b200 import _lib
    return _lib.load()


def test_library_exports_every_declared_symbol(lib):
    from hvit_b200 import _lib
    header = open(os.path.join(ROOT, "include", "hvit.h")).read()
    declared = set(re.findall(r"\b(hvit_[a-z0-9_]+)\s*\(", header))
    assert declared, "no declarations parsed"
    assert declared == set(_lib.SYMBOLS), declared ^ set(_lib.SYMBOLS)
    for name in declared:
        assert hasattr(lib, name), name
    assert lib.hvit_version() >= 100


def test_struct_layout_matches_header(lib):
    """hvit_workspace_bytes parses the ctypes ModelCfg; a layout drift shows up as nonsense sizes / errors."""
    from hvit_b200.models import HybridViT
    m = HybridViT()
    for prec, es in ((0, 4), (1, 2), (2, 2)):
        cfg = m._c_cfg(prec)
        n1 = lib.hvit_workspace_bytes(C.byref(cfg), 1, 257, 501, 0)
        n64 = lib.hvit_workspace_bytes(C.byref(cfg), 64, 257, 501, 64000)
        assert n1 > 0 and n64 > 60 * n1
        # stem output [B,128,250,64] alone
        assert n1 > 128 * 250 * 64 * es
    cfg = m._c_cfg(2)
    assert lib.hvit_workspace_bytes(C.byref(cfg), 1, 257, 8, 0) == 0          # no patches
    assert b"patch" in lib.hvit_last_error()
    assert lib.hvit_workspace_bytes(C.byref(cfg), 1, 257, 501, 63999) == 0    # T inconsistent with n_samples
    cfg.embed_dim = 500
    assert lib.hvit_workspace_bytes(C.byref(cfg), 1, 257, 501, 0) == 0
    assert b"head_dim" in lib.hvit_last_error()


def test_state_dict_keys_match_reference_spec(oracle):
    from hvit_b200.models import HybridViT
    for over in ({}, dict(encoder_channels=[64, 64, 128], embed_dim=128, num_heads=2, num_layers=2,
                          decoder_channels=[128, 64, 64, 1])):
        cfg = oracle.full_cfg(over)
        m = HybridViT(**{k: cfg[k] for k in ("encoder_channels", "embed_dim", "num_heads", "num_layers",
                                             "decoder_channels")})
        spec = oracle.state_dict_spec(cfg)
        sd = m.state_dict()
        assert list(sd.keys()) == [k for k, _, _ in spec]
        for k, shape, _ in spec:
            assert tuple(sd[k].shape) == tuple(shape), k
        m.load_state_dict(oracle.make_state_dict(cfg, seed=1), strict=True)


def test_literal_init_is_bit_identical_to_reference(oracle, golden):
    """Same torch.manual_seed -> same parameters as the reference's HybridViT() (digest recorded by make_golden.py)."""
    from hvit_b200.models import HybridViT
    _, meta = golden
    torch.manual_seed(0)
    m = HybridViT()
    assert oracle.state_dict_digest(m.state_dict()) == meta["literal_init_seed0"]["weights_sha256"]
    assert m.count_parameters()["total"] == 28454976


def test_oracle_on_literal_init_matches_reference_output(oracle, golden):
    from hvit_b200.models import HybridViT
    arrays, _ = golden
    torch.manual_seed(0)
    sd = HybridViT().state_dict()
    x = torch.rand(2, 1, 257, 63, generator=torch.Generator().manual_seed(5))
    with torch.no_grad():
        y = oracle.hybrid_vit_forward(sd, x)
    assert oracle.max_rel_err(y.numpy(), arrays["literal_init_seed0/model_out"]) <= 1e-5


def test_create_from_yaml_and_quirks():
    from hvit_b200.utils.config import load_all_configs, merge_configs
    from hvit_b200.models import create_hybrid_vit
    cfg = load_all_configs(os.path.join(ROOT, "config"))
    assert cfg["audio"]["n_fft"] == 512 and cfg["data"]["hop_length"] == 128
    m = create_hybrid_vit(cfg)
    assert m.arch["embed_dim"] == 512 and m.arch["num_layers"] == 6 and m.arch["encoder_pool_sizes"] == [2, 2, 1]
    assert create_hybrid_vit(None).arch == m.arch           # defaults == shipped YAML
    assert merge_configs({"a": {"b": 1, "c": 2}}, {"a": {"b": 3}}) == {"a": {"b": 3, "c": 2}}
    with pytest.raises(FileNotFoundError):
        load_all_configs(os.path.join(ROOT, "no_such_dir"))


def test_error_conventions():
    from hvit_b200.models import HybridViT, ConvBlock, MultiHeadSelfAttention
    from hvit_b200.inference import AudioEnhancer
    from hvit_b200.utils.checkpoint import load_model_weights
    with pytest.raises(ValueError):
        ConvBlock(1, 8, activation="swish")                 # reference components.py:77
    with pytest.raises(AssertionError):
        MultiHeadSelfAttention(100, num_heads=8)            # reference attention.py:46
    with pytest.raises(FileNotFoundError):
        load_model_weights("/no/such/ckpt.pt", HybridViT(num_layers=1))
    tiny = HybridViT(encoder_channels=[64, 64, 128], embed_dim=128, num_heads=2, num_layers=1,
                     decoder_channels=[128, 64, 64, 1])
    with pytest.raises(RuntimeError):
        AudioEnhancer(tiny, device="cpu")                   # no CPU path
    with pytest.raises(RuntimeError):
        tiny.eval()(torch.rand(1, 1, 257, 63))              # CPU tensor
    with pytest.raises(RuntimeError):
        tiny.encoder[0](torch.rand(1, 1, 8, 8))             # blocks are parameter containers


def test_checkpoint_roundtrip(tmp_path, oracle):
    from hvit_b200.models import HybridViT
    from hvit_b200.utils.checkpoint import load_model_weights
    kw = dict(encoder_channels=[64, 64, 128], embed_dim=128, num_heads=2, num_layers=1, decoder_channels=[128, 64, 64, 1])
    a, b = HybridViT(**kw), HybridViT(**kw)
    torch.save({"model_state_dict": a.state_dict()}, tmp_path / "wrapped.pt")
    torch.save(a.state_dict(), tmp_path / "bare.pt")
    for f in ("wrapped.pt", "bare.pt"):
        load_model_weights(str(tmp_path / f), b, device="cpu", strict=True)
        assert oracle.state_dict_digest(a.state_dict()) == oracle.state_dict_digest(b.state_dict())


def test_weight_packing_cpu(oracle):
    """BN folding and the x2-upsample parity kernels are exact re-expressions (checked in fp64 on the CPU)."""
    import torch.nn.functional as F
    from hvit_b200.models.packing import fold_bn, up2_parity_kernels, conv_khwc
    g = torch.Generator().manual_seed(0)
    sd = {"bn.weight": torch.rand(8, generator=g) + 0.5, "bn.bias": torch.randn(8, generator=g),
          "bn.running_mean": torch.randn(8, generator=g), "bn.running_var": torch.rand(8, generator=g) + 0.5}
    x = torch.randn(2, 8, 5, 5, generator=g)
    scale, shift = fold_bn(sd, "bn")
    ref = F.batch_norm(x, sd["bn.running_mean"], sd["bn.running_var"], sd["bn.weight"], sd["bn.bias"], False, 0.0, 1e-5)
    assert torch.allclose(x * scale[None, :, None, None] + shift[None, :, None, None], ref, atol=1e-5)
    w = torch.randn(6, 4, 3, 3, generator=g)
    k = up2_parity_kernels(w).double()
    xi = torch.randn(1, 4, 7, 5, generator=g).double()
    ref = F.conv2d(F.interpolate(xi, scale_factor=2, mode="nearest"), w.double(), padding=1)
    xp = F.pad(xi, (1, 1, 1, 1))
    out = torch.zeros_like(ref)
    for py in range(2):
        for px in range(2):
            out[:, :, py::2, px::2] = F.conv2d(xp[:, :, py:py + 8, px:px + 6], k[py * 2 + px].permute(0, 3, 1, 2))
    assert float((out - ref).abs().max()) < 1e-5
    assert conv_khwc(w).shape == (6, 3, 3, 4)


def test_wav_io_roundtrip(tmp_path):
    """load_audio / save_audio keep the reference's signatures (utils/audio_processing.py:15-64) and cover the subtypes
    soundfile would write: PCM_16 (default), PCM_24, PCM_32, FLOAT; stereo -> mono is the channel mean; a sample-rate
    mismatch raises unless a resampler is named (librosa's default soxr_hq is unavailable)."""
    import wave
    from hvit_b200.utils.audio_processing import load_audio, save_audio
    x = (0.5 * np.sin(np.arange(8000) * 0.05)).astype(np.float32)
    for subtype, tol in (("PCM_16", 2.0 ** -15), ("PCM_24", 2.0 ** -23), ("PCM_32", 1e-7), ("FLOAT", 0.0)):
        save_audio(tmp_path / f"{subtype}.wav", x, 16000, subtype=subtype)
        y, sr = load_audio(tmp_path / f"{subtype}.wav", sr=16000)
        assert sr == 16000 and y.shape == x.shape and y.dtype == np.float32
        assert np.abs(x - y).max() <= tol, subtype
    with wave.open(str(tmp_path / "PCM_16.wav"), "rb") as f:       # readable by an independent parser, 16-bit mono
        assert (f.getnchannels(), f.getsampwidth(), f.getframerate(), f.getnframes()) == (1, 2, 16000, 8000)
        pcm = np.frombuffer(f.readframes(8000), dtype="<i2")
    assert np.array_equal(pcm, np.rint(x.astype(np.float64) * 32767.0).astype(np.int16))   # libsndfile's float -> short
    st = np.stack([x, -0.5 * x], axis=1)
    save_audio(tmp_path / "st.wav", st, 16000, subtype="FLOAT")
    m, _ = load_audio(tmp_path / "st.wav", sr=16000, mono=True)
    assert np.allclose(m, 0.25 * x, atol=1e-7)
    s2, _ = load_audio(tmp_path / "st.wav", sr=16000, mono=False)
    assert s2.shape == (2, 8000)
    seg, _ = load_audio(tmp_path / "FLOAT.wav", sr=16000, offset=0.1, duration=0.2)
    assert np.array_equal(seg, x[1600:4800])
    save_audio(tmp_path / "r8k.wav", x, 8000, subtype="FLOAT")
    with pytest.raises(NotImplementedError):
        load_audio(tmp_path / "r8k.wav", sr=16000)
    up, sr = load_audio(tmp_path / "r8k.wav", sr=16000, res_type="polyphase")
    from scipy.signal import resample_poly
    assert sr == 16000 and np.allclose(up, resample_poly(x, 2, 1).astype(np.float32), atol=1e-6)
    with pytest.raises(ValueError):
        save_audio(tmp_path / "bad.wav", x, 16000, subtype="ULAW")


def test_shard_range():
    from hvit_b200.inference.sharding import shard_range
    for n in (0, 1, 7, 64, 512, 513):
        for world in (1, 2, 3, 8):
            spans = [shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_range(4, 2, 2)


def _dp_worker(rank, world, port, q):
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    from hvit_b200.inference.sharding import enhance_sharded
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    clips = torch.arange(7 * 5, dtype=torch.float32).reshape(7, 5)
    seen = []

    def fake_enhance(x):            # stands in for AudioEnhancer.enhance_device (no GPU here)
        seen.append(x.shape[0])
        return x * 2 + 1

    full = enhance_sharded(clips, fake_enhance, gather=True, micro_batch=3)
    mine = enhance_sharded(clips, fake_enhance, gather=False, micro_batch=3)
    q.put((rank, full.tolist(), mine.shape[0], seen))
    dist.destroy_process_group()


def test_data_parallel_sharding_gloo_world2():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_dp_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    expect = (torch.arange(35, dtype=torch.float32).reshape(7, 5) * 2 + 1).tolist()
    for rank, full, n_mine, seen in res:
        assert full == expect
        assert n_mine == (4 if rank == 0 else 3)
        assert max(seen) <= 3
