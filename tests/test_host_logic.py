"""CPU-only checks of the host-side mirror of the reference interface and of the C-ABI library."""
import ctypes
import os
import re
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_registries_match_reference_names():
    from storm_b200 import BackboneRegistry, CorrectorRegistry, PredictorRegistry, SDERegistry
    assert {"ncsnpp", "ncsnpplarge", "ncsnpp12M", "ncsnpp6M"} <= set(BackboneRegistry.get_all_names())
    assert "ncsnpp-large" in BackboneRegistry.get_all_names()          # README.md:43 spelling
    assert SDERegistry.get_all_names() == ["ouve"]
    assert set(PredictorRegistry.get_all_names()) == {"reverse_diffusion", "euler_maruyama", "none"}
    assert set(CorrectorRegistry.get_all_names()) == {"ald", "langevin", "none"}
    with pytest.raises(ValueError, match="unknown"):
        BackboneRegistry.get_by_name("resnet")
    with pytest.warns(UserWarning):
        BackboneRegistry.register("ncsnpp6M")(BackboneRegistry.get_by_name("ncsnpp6M"))


def test_backbone_state_dict_contract():
    from storm_b200 import BackboneRegistry, ncsnpp_spec
    m = BackboneRegistry.get_by_name("ncsnpp")(nf=8, unknown_kwarg=1, gpus=0)     # swallows unknown kwargs (ncsnpp.py:65)
    sd = m.state_dict()
    assert list(sd.keys()) == [k for k, _ in ncsnpp_spec.param_table(m.cfg)]
    assert not m.FORCE_STFT_OUT
    with pytest.raises(RuntimeError):
        m.load_state_dict({k: v for k, v in list(sd.items())[:-1]}, strict=True)
    with pytest.raises(NotImplementedError):
        BackboneRegistry.get_by_name("ncsnpp")(nf=8, resblock_type="ddpm")
    large = BackboneRegistry.get_by_name("ncsnpplarge")
    assert ncsnpp_spec.param_count(ncsnpp_spec.config_from_kwargs(dict(large.DEFAULTS), {})) == 65590822
    # default init_scale = 0 reproduces the reference's degenerate 1e-10 variance scale (layers.py:90)
    m0 = BackboneRegistry.get_by_name("ncsnpp")(nf=8)
    assert float(m0.state_dict()["all_modules.4.Conv_1.weight"].abs().max()) < 1e-4
    assert float(m0.state_dict()["all_modules.4.Conv_0.weight"].abs().max()) > 1e-3


def test_no_gpu_fails_loudly():
    from storm_b200 import BackboneRegistry, ScoreModel
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    m = BackboneRegistry.get_by_name("ncsnpp")(nf=8)
    with pytest.raises(RuntimeError, match="no CUDA device|CUDA"):
        m(torch.zeros(1, 2, 8, 8, dtype=torch.complex64), torch.ones(1))
    with pytest.raises(RuntimeError, match="CUDA"):
        ScoreModel(backbone="ncsnpp", nf=8).enhance(torch.zeros(1, 4000))


def test_library_exports_every_declared_symbol():
    from storm_b200 import _native
    lib = _native.lib()
    hdr = open(os.path.join(ROOT, "include", "storm_b200.h")).read()
    declared = set(re.findall(r"\b(storm_[a-z_0-9]+)\s*\(", hdr))
    assert declared, "no declarations parsed"
    for sym in declared:
        assert hasattr(lib, sym), f"{sym} declared in include/storm_b200.h but not exported"
    assert declared == set(_native.exported_symbols())
    assert lib.storm_version().decode().startswith("storm_b200")
    # handle-free host helpers of the ABI (index work, bit-exact): frame count and pad_spec amounts
    sp = _native.make_spec_config()
    for L, frames, padded in [(32640, 256, 256), (64000, 501, 512), (128000, 1001, 1024), (5000, 40, 64), (8192, 65, 128)]:
        assert _native.num_frames(L, sp) == frames == 1 + L // 128
        assert _native.padded_frames(frames) == padded
    sc = _native.make_sampler_config("reverse_diffusion", "ald", 30, 1)
    assert lib.storm_pc_num_draws(ctypes.byref(sc)) == 61
    with pytest.raises(ValueError):
        _native.make_sampler_config("nope", "ald")


def test_pad_spec_and_sde_formulas():
    from storm_b200.sdes import OUVESDE
    from storm_b200.util.other import pad_spec, padded_frames
    from oracle import storm_oracle as O
    Y = torch.randn(1, 1, 256, 501, dtype=torch.complex64)
    P = pad_spec(Y)
    assert P.shape[-1] == 512 == padded_frames(501) and torch.equal(P[..., :501], Y) and torch.count_nonzero(P[..., 501:]) == 0
    assert pad_spec(torch.zeros(1, 1, 4, 64)).shape[-1] == 64
    sde, ref = OUVESDE(theta=1.5, sigma_min=0.05, sigma_max=0.5, N=30), O.OUVE()
    t = torch.linspace(1.0, 0.03, 30)
    np.testing.assert_allclose(sde._std(t).numpy(), ref.std(t).numpy(), rtol=1e-6)
    x, y = torch.randn(2, 1, 4, 4, dtype=torch.complex64), torch.randn(2, 1, 4, 4, dtype=torch.complex64)
    drift, diff = sde.sde(x, t[:2], y)
    assert torch.allclose(drift, 1.5 * (y - x)) and torch.allclose(diff, ref.diffusion(t[:2]))
    assert sde.copy().N == 30 and sde.T == 1


def test_python_level_sampler_matches_oracle_with_custom_score_fn():
    """Third-party code written against the reference sampler API (custom score_fn) keeps working."""
    from storm_b200 import sampling
    from storm_b200.sdes import OUVESDE
    from oracle import storm_oracle as O
    g = torch.Generator().manual_seed(0)
    y = torch.complex(torch.randn(2, 1, 8, 8, generator=g), torch.randn(2, 1, 8, 8, generator=g))
    score = lambda x, t, yy: -(x - yy) * t[:, None, None, None]        # noqa: E731
    draws = [torch.complex(torch.randn(2, 1, 8, 8, generator=g), torch.randn(2, 1, 8, 8, generator=g)) * 0.7071 for _ in range(7)]
    it = iter(draws)
    orig = torch.randn_like
    torch.randn_like = lambda x, **kw: next(it)
    try:
        x, nfe = sampling.get_pc_sampler("reverse_diffusion", "ald", OUVESDE(N=3), score, y, snr=0.5, eps=0.03)()
    finally:
        torch.randn_like = orig
    it2 = iter(draws)
    xo, nfeo = O.pc_sample(score, y, O.OUVE(), 3, noise=lambda x: next(it2))
    assert nfe == nfeo == 6 and torch.allclose(x, xo, rtol=1e-5, atol=1e-6)


def test_checkpoint_loader_reads_lightning_format(tmp_path):
    from storm_b200 import ScoreModel
    src = ScoreModel(backbone="ncsnpp", sde="ouve", nf=8, init_scale=1.0)
    shadow = [p.detach() * 0.5 for p in src.dnn.parameters()]
    ckpt = dict(state_dict=src.state_dict(), hyper_parameters=dict(backbone="ncsnpp", sde="ouve", nf=8, t_eps=0.03, theta=1.5),
                ema=dict(decay=0.999, num_updates=1, shadow_params=shadow, collected_params=None))
    path = tmp_path / "m.ckpt"
    torch.save(ckpt, path)
    m = ScoreModel.load_from_checkpoint(str(path), base_dir="", batch_size=1, num_workers=0, kwargs=dict(gpu=False))
    k = "all_modules.4.Conv_0.weight"
    assert torch.equal(m.dnn.state_dict()[k], src.dnn.state_dict()[k])
    m.eval(no_ema=False)                                     # EMA shadow weights swapped in (model.py:97-108)
    assert torch.equal(m.dnn.state_dict()[k], shadow[[n for n, _ in m.dnn.named_parameters()].index(k)])
    m.train(True)
    assert torch.equal(m.dnn.state_dict()[k], src.dnn.state_dict()[k])


def test_cli_flags_match_reference():
    sys.path.insert(0, ROOT)
    import enhancement
    p = enhancement.build_parser()
    a = p.parse_args(["--test_dir", "a", "--enhanced_dir", "b", "--ckpt", "c", "--mode", "score-only"])
    assert (a.corrector, a.corrector_steps, a.snr, a.N) == ("ald", 1, 0.5, 50)      # enhancement.py:33-36 defaults
    assert enhancement.shard(list(range(10)), 1, 4) == [1, 5, 9]


def test_bucket_utterances_groups_by_padded_frames():
    from storm_b200.parallel import bucket_utterances, padded_frame_count
    # 32640 -> 256 frames -> 256; 32000 -> 251 -> 256; 33000 -> 258 -> 320; 8000 -> 63 -> 64; 8191 -> 64 -> 64; 8192 -> 65 -> 128
    lens = [32640, 33000, 32000, 8000, 8192, 8191, 32100, 32200]
    assert [padded_frame_count(L) for L in lens] == [256, 320, 256, 64, 128, 64, 256, 256]
    b = bucket_utterances(lens, max_batch=3)
    assert b == [[3, 5], [4], [0, 2, 6], [7], [1]]
    assert sorted(i for g in b for i in g) == list(range(len(lens)))
    with pytest.raises(ValueError):
        bucket_utterances([100, 0])
    with pytest.raises(ValueError):
        bucket_utterances([100], max_batch=0)


def test_bucket_matches_reference_pad_spec_rule():
    """Frame count and padding of every length in a range equal torch.stft + pad_spec of the reference path (fixture shapes)."""
    from conftest import load_golden
    from storm_b200.parallel import padded_frame_count
    g = load_golden("stft")
    for tag in ("a", "b", "c"):
        L = g[f"{tag}/sig"].shape[-1]
        assert padded_frame_count(L) == int(g[f"{tag}/padded_T"])
        assert 1 + L // 128 == g[f"{tag}/stft"].shape[-1]


def test_wavio_roundtrip_and_pipeline(tmp_path):
    from storm_b200.util import wavio
    g = torch.Generator().manual_seed(0)
    paths = []
    for i, L in enumerate([1600, 2400, 800]):
        x = (torch.rand(1, L, generator=g) - 0.5) * 0.8
        p = str(tmp_path / f"u{i}.wav")
        wavio.save_wav(p, x, 16000)
        y, sr = wavio.load_wav(p)
        assert sr == 16000 and y.shape == (1, L) and torch.allclose(x, y, atol=1e-6)
        assert wavio.wav_length(p) == L
        paths.append(p)
    # 16-bit PCM through the header-only length probe and the scipy reader
    from scipy.io import wavfile
    pcm = (np.sin(np.arange(1000) * 0.05) * 20000).astype(np.int16)
    p16 = str(tmp_path / "pcm16.wav")
    wavfile.write(p16, 16000, pcm)
    assert wavio.wav_length(p16) == 1000
    y, sr = wavio._scipy_load(p16)
    assert y.shape == (1, 1000) and abs(float(y[0, 10]) - pcm[10] / 32768.0) < 1e-7
    # prefetcher keeps order, async writer flushes on close
    got = list(wavio.Prefetcher(paths, lambda q: wavio.load_wav(q)[0].shape[1], depth=2, workers=2))
    assert got == [1600, 2400, 800]
    with wavio.AsyncWriter(workers=2) as w:
        for i, q in enumerate(paths):
            w.save(str(tmp_path / f"o{i}.wav"), wavio.load_wav(q)[0] * 0.5, 16000)
    assert all((tmp_path / f"o{i}.wav").exists() for i in range(3))


def test_ragged_length_validation_is_host_side():
    """The checks in front of storm_*_ragged (the library cannot read device lengths) - no GPU needed."""
    from storm_b200 import _native
    sp = _native.make_spec_config()
    assert _native.check_lengths([5000, 4000, 7000, 8100], 4, 8100, sp) == [5000, 4000, 7000, 8100]
    assert _native.check_lengths(torch.tensor([32640, 32000]), 2, 32640, sp) == [32640, 32000]
    for bad, B, L in [([5000, 4000], 3, 8000),            # wrong count
                      ([5000, 9000], 2, 8100),            # longer than the row
                      ([5000, 0], 2, 8100),               # empty
                      ([5000, 200], 2, 8100),             # shorter than the reflect padding (255)
                      ([5000, 9000], 2, 9000)]:           # 40 frames -> 64 vs 71 frames -> 128: different buckets
        with pytest.raises(ValueError):
            _native.check_lengths(bad, B, L, sp)
    assert _native.check_lengths([5000, 200], 2, 8100) == [5000, 200]     # without a SpecConfig only the range is checked


def test_parameter_order_matches_the_reference_modules():
    """tests/golden/param_order.json = named_parameters() of the reference NCSNpp (oracle/gen_golden.py::gen_param_order):
    the order torch_ema's shadow list follows, and the only frozen parameter (the Fourier W)."""
    import json
    from storm_b200 import BackboneRegistry
    from storm_b200.model import FROZEN_PARAMS
    ref = json.load(open(os.path.join(ROOT, "tests", "golden", "param_order.json")))
    for tag, kw in [("ncsnpp_4ch", dict(input_channels=4)), ("ncsnpp_6ch", dict(input_channels=6)),
                    ("ncsnpp_disc", dict(input_channels=2, discriminative=True))]:
        m = BackboneRegistry.get_by_name("ncsnpp")(nf=8, **kw)
        mine = [(n, list(p.shape)) for n, p in m.named_parameters()]
        assert mine == [(n, shp) for n, shp, _ in ref[tag]], tag
        assert [n for n, _, rg in ref[tag] if not rg] == list(FROZEN_PARAMS)


def test_ema_shadow_mapping_is_name_and_shape_checked():
    from storm_b200 import BackboneRegistry
    from storm_b200.model import map_ema_shadow
    m = BackboneRegistry.get_by_name("ncsnpp")(nf=8)
    named = list(m.named_parameters())
    full = [torch.full_like(p, float(i)) for i, (_, p) in enumerate(named)]
    pairs = map_ema_shadow(named, full)                                   # torch_ema 0.3: every parameter has a shadow
    assert [n for n, _ in pairs] == [n for n, _ in named]
    trainable = [t for (n, _), t in zip(named, full) if n != "all_modules.0.W"]
    pairs = map_ema_shadow(named, trainable)                              # torch_ema 0.2: frozen parameters have none
    assert "all_modules.0.W" not in [n for n, _ in pairs] and len(pairs) == len(named) - 1
    assert all(float(sh.flatten()[0]) == float([i for i, (k, _) in enumerate(named) if k == n][0]) for n, sh in pairs)
    with pytest.raises(RuntimeError, match="cannot align"):
        map_ema_shadow(named, full[:-3])
    bad = list(full)
    bad[5] = torch.zeros(3, 3)
    with pytest.raises(RuntimeError, match="shape"):
        map_ema_shadow(named, bad)
