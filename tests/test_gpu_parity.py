"""Parity of the CUDA path (through the C ABI) with the reference-generated fixtures and the CPU oracle.
Run on the B200 box:  python -m pytest tests -m gpu"""
import numpy as np
import pytest
import torch

from conftest import load_golden, rel_l2, sub_sd
from oracle import storm_oracle as O

pytestmark = pytest.mark.gpu
T = torch.from_numpy
TOL = 1e-3          # north-star tolerance on fp32 spectrogram values
TIGHT = 2e-4        # what the fp32 / bf16x3 kernels are expected to reach


def _model(name="ncsnpp", sd=None, **kw):
    from storm_b200 import BackboneRegistry
    m = BackboneRegistry.get_by_name(name)(**kw)
    if sd is not None:
        m.load_state_dict(sd, strict=True)
    return m.cuda()


@pytest.mark.parametrize("fix,kw", [
    ("ncsnpp_tiny", dict(nf=8)),
    ("ncsnpp_tiny_6ch", dict(nf=8, input_channels=6)),
    ("ncsnpp_tiny_disc", dict(nf=8, discriminative=True)),
    ("ncsnpplarge_tiny", dict(nf=8, ch_mult=(1, 1, 2, 2, 2, 2, 2), num_res_blocks=2, attn_resolutions=(16,))),
])
def test_tiny_networks_vs_reference(fix, kw):
    g = load_golden(fix)
    m = _model("ncsnpp", sub_sd(g, "sd/"), **kw)
    t = T(g["t"]).cuda() if "t" in g else None
    y = m(T(g["x"]).cuda(), t).cpu()
    assert y.shape == g["y"].shape
    assert rel_l2(y, g["y"]) < TIGHT


def _full(name):
    from storm_b200 import ncsnpp_spec
    fix = {"ncsnpp": "ncsnpp_full_64x64", "ncsnpplarge": "ncsnpplarge_full_256x64"}[name]
    g = load_golden(fix)
    m = _model(name)
    sd = ncsnpp_spec.random_state_dict(m.cfg, seed=int(g["seed"]), init_scale=float(g["init_scale"]), bias_std=float(g["bias_std"]))
    m.load_state_dict(sd)
    return g, m, sd


@pytest.mark.parametrize("name", ["ncsnpp", "ncsnpplarge"])
def test_full_width_vs_reference(name):
    g, m, _ = _full(name)
    y = m(T(g["x"]).cuda(), T(g["t"]).cuda()).cpu()
    err = rel_l2(y, g["y"])
    print(f"{name}: rel-L2 vs reference = {err:.3e}")
    assert err < TOL


def test_layer_taps_vs_oracle():
    """Every module output of the full-width ncsnpp against the oracle's activation at the same point."""
    g, m, sd = _full("ncsnpp")
    eng = m.engine()
    eng.set_taps(True)
    y = m(T(g["x"]).cuda(), T(g["t"]).cuda())
    taps = {}
    O.ncsnpp_forward(sd, O.NetCfg(), T(g["x"]), T(g["t"]), taps=taps)
    worst = 0.0
    for name, ref in taps.items():
        got = eng.get_tap(name).cpu()
        assert got.shape == ref.shape, name
        e = rel_l2(got, ref)
        worst = max(worst, e)
        print(f"{name:10s} {tuple(ref.shape)} rel-L2 {e:.3e}")
        assert e < TOL, f"layer {name} deviates: {e}"
    eng.set_taps(False)
    print("worst layer error", worst)


def _noise(arr):
    return torch.from_numpy(np.ascontiguousarray(arr)).cuda()


@pytest.mark.parametrize("tag,pred,corr", [("rd_ald", "reverse_diffusion", "ald"), ("rd_ald2", "reverse_diffusion", "ald"),
                                           ("em_langevin", "euler_maruyama", "langevin"),
                                           ("rd_none", "reverse_diffusion", "none"), ("none_ald", "none", "ald")])
def test_pc_sampler_vs_reference(tag, pred, corr):
    from storm_b200 import _native
    g = load_golden("sampler")
    m = _model("ncsnpp", sub_sd(g, "sd/"), nf=8)
    sc = _native.make_sampler_config(pred, corr, int(g[f"{tag}/N"]), int(g[f"{tag}/csteps"]), float(g[f"{tag}/snr"]), 0.03)
    x, nfe = m.engine().pc_sample(T(g["y"]).cuda(), sc, noise=_noise(g[f"{tag}/noise"]))
    assert nfe == int(g[f"{tag}/nfe"])
    assert rel_l2(x.cpu(), g[f"{tag}/x"]) < TIGHT


def test_registry_level_sampler_and_errors():
    from storm_b200 import ScoreModel, sampling
    g = load_golden("sampler")
    model = ScoreModel(backbone="ncsnpp", sde="ouve", nf=8)
    model.load_state_dict({"dnn." + k: v for k, v in sub_sd(g, "sd/").items()})
    model.cuda()
    Y = T(g["y"]).cuda()
    sampler = model.get_pc_sampler("reverse_diffusion", "ald", Y, N=int(g["rd_ald/N"]), snr=0.5, noise=_noise(g["rd_ald/noise"]))
    x, nfe = sampler()
    assert nfe == 6 and rel_l2(x.cpu(), g["rd_ald/x"]) < TIGHT
    with pytest.raises(ValueError):
        model.get_pc_sampler("no_such_predictor", "ald", Y, N=3)
    with pytest.raises(ValueError):
        sampling.CorrectorRegistry.get_by_name("nope")
    # score function contract: ScoreModel.forward = -dnn(cat[x, y], t)
    t = torch.tensor([0.5, 0.7], device="cuda")
    s = model(Y, t, Y)
    sref = O.make_score_fn(sub_sd(g, "sd/"), O.NetCfg(nf=8))(T(g["y"]), t.cpu(), T(g["y"]))
    assert rel_l2(s.cpu(), sref) < TIGHT


def test_ode_sampler_vs_reference():
    from storm_b200 import ScoreModel
    g = load_golden("sampler")
    model = ScoreModel(backbone="ncsnpp", sde="ouve", nf=8)
    model.load_state_dict({"dnn." + k: v for k, v in sub_sd(g, "sd/").items()})
    model.cuda()
    Y = T(g["y"]).cuda()
    sampler = model.get_ode_sampler(Y, N=int(g["ode/N"]), minibatch=None, noise=_noise(g["ode/noise"]))
    x, nfe = sampler()
    assert abs(nfe - int(g["ode/nfe"])) <= 12          # adaptive solver: step count may move by a stage or two
    assert rel_l2(x.cpu(), g["ode/x"]) < 1e-3


@pytest.mark.parametrize("tag", ["a", "b", "c"])
def test_stft_istft_vs_reference(tag):
    from storm_b200 import _native, SpecsDataModule
    g = load_golden("stft")
    dm = SpecsDataModule()
    sig = T(g[f"{tag}/sig"]).cuda()
    S = dm.stft(sig)
    assert tuple(S.shape) == g[f"{tag}/stft"].shape                      # frame/bin counts: exact
    assert rel_l2(S.cpu(), g[f"{tag}/stft"]) < 1e-5
    L = sig.shape[-1]
    T_pad = _native.padded_frames(_native.num_frames(L, dm.spec_config()))
    assert T_pad == int(g[f"{tag}/padded_T"])
    Sf = _native.stft(sig, None, dm.spec_config(), T_pad)
    frames = g[f"{tag}/fwd"].shape[-1]
    assert rel_l2(Sf[..., :frames].cpu(), g[f"{tag}/fwd"]) < 1e-5
    assert torch.count_nonzero(Sf[..., frames:]) == 0                    # pad_spec zeros: exact
    back = dm.istft(dm.spec_back(T(g[f"{tag}/fwd"]).cuda()), L)
    assert tuple(back.shape) == g[f"{tag}/back"].shape
    assert rel_l2(back.cpu(), g[f"{tag}/back"]) < 1e-5
    back2 = m_istft_fused(T(g[f"{tag}/fwd"]).cuda(), L, dm)
    assert rel_l2(back2.cpu(), g[f"{tag}/back"]) < 1e-5


def m_istft_fused(spec, L, dm):
    from storm_b200.data_module import _scratch_engine
    return _scratch_engine().istft(spec, None, L, dm.spec_config())


def test_stft_sqrthann_identity_transform():
    from storm_b200 import SpecsDataModule
    g = load_golden("stft")
    dm = SpecsDataModule(spec_factor=0.33, spec_abs_exponent=1.0, window="sqrthann")
    S = dm.stft(T(g["d/sig"]).cuda())
    assert rel_l2(S.cpu(), g["d/stft"]) < 1e-5
    assert rel_l2(dm.spec_fwd(S).cpu(), g["d/fwd"]) < 1e-5
    assert rel_l2(dm.istft(dm.spec_back(dm.spec_fwd(S)), 3000).cpu(), g["d/back"]) < 1e-5


def test_enhance_chain_vs_reference():
    """normalise -> STFT -> transform -> pad -> PC sampler (injected noise) -> inverse, as ScoreModel.enhance."""
    from storm_b200 import _native
    g = load_golden("enhance")
    m = _model("ncsnpp", sub_sd(g, "sd/"), nf=8)
    sp = _native.make_spec_config()
    y = T(g["y"]).cuda()
    L = y.shape[1]
    norm = _native.wave_norm(y)
    assert float(norm[0]) == float(np.abs(g["y"]).max())                  # exact
    T_pad = _native.padded_frames(_native.num_frames(L, sp))
    Y = _native.stft(y, norm, sp, T_pad).unsqueeze(1)
    assert tuple(Y.shape) == g["Y"].shape
    assert rel_l2(Y.cpu(), g["Y"]) < 1e-5
    sc = _native.make_sampler_config("reverse_diffusion", "ald", int(g["N"]), 1, 0.5, 0.03)
    X, nfe = m.engine().pc_sample(Y, sc, noise=_noise(g["noise"]))
    assert nfe == int(g["nfe"])
    assert rel_l2(X.cpu(), g["sample"]) < TIGHT
    x_hat = m.engine().istft(X.squeeze(1), norm, L, sp)
    assert rel_l2(x_hat.squeeze().cpu(), g["x_hat"]) < TIGHT


def test_philox_sampler_matches_oracle_and_is_sharding_invariant():
    from storm_b200 import _native
    g = load_golden("sampler")
    sd = sub_sd(g, "sd/")
    m = _model("ncsnpp", sd, nf=8)
    eng = m.engine()
    Y = T(g["y"]).cuda()
    B, n = Y.shape[0], Y[0].numel()
    sc = _native.make_sampler_config("reverse_diffusion", "ald", 2, 1, 0.5, 0.03)
    ids = [7, 11]
    X, _ = eng.pc_sample(Y, sc, seed=123, utt_ids=ids)
    nd = 1 + 2 * 2
    draws = torch.stack([torch.stack([_native.philox_normal(123, u, d, n).reshape(Y.shape[1:]) for u in ids]) for d in range(nd)])
    z = draws.reshape(nd * B, -1)
    assert abs(float(z.real.var()) - 0.5) < 0.02 and abs(float(z.imag.var()) - 0.5) < 0.02 and abs(float(z.real.mean())) < 0.01
    it = iter([d.cpu() for d in draws])
    xo, _ = O.pc_sample(O.make_score_fn(sd, O.NetCfg(nf=8)), T(g["y"]), O.OUVE(), 2, noise=lambda x: next(it))
    assert rel_l2(X.cpu(), xo) < TIGHT
    # each utterance alone (another batch composition / another "rank") gives bit-identical results
    for b in range(B):
        Xb, _ = eng.pc_sample(Y[b:b + 1], sc, seed=123, utt_ids=[ids[b]])
        assert torch.equal(Xb[0], X[b])


def test_enhance_batch_api_and_shapes():
    from storm_b200 import ScoreModel
    model = ScoreModel(backbone="ncsnpp", sde="ouve", nf=8, init_scale=1.0).cuda()
    ys = torch.stack([O.synth_utterance(b, 5000) for b in range(3)])
    out = model.enhance_batch(ys, N=2, seed=5, utt_ids=[0, 1, 2])
    assert out.shape == ys.shape and out.device == ys.device and torch.isfinite(out).all()
    one = model.enhance_batch(ys[1:2], N=2, seed=5, utt_ids=[1])
    assert torch.equal(one[0], out[1])                                   # batch-composition invariance
    x1 = model.enhance(ys[0:1], N=2)
    assert x1.shape == (5000,) and x1.device.type == "cpu"
    x2, nfe, rtf = model.enhance(ys[0:1], N=2, timeit=True)
    assert nfe == 4 and rtf > 0
    with pytest.raises(ValueError):
        model.enhance(ys[0:1], sampler_type="bogus", return_stft=True)


def test_full_size_properties():
    """BASELINE shape (256x256, full-width ncsnpp): finite output, batch-order equivariance (bit-exact),
    and the score network's exact scaling law in the output layer (d(t) * t is affine in the pyramid)."""
    from storm_b200 import ncsnpp_spec
    m = _model("ncsnpp")
    m.load_state_dict(ncsnpp_spec.random_state_dict(m.cfg, seed=0, init_scale=1.0, bias_std=0.05))
    gen = torch.Generator().manual_seed(0)
    x = torch.complex(torch.randn(2, 2, 256, 256, generator=gen), torch.randn(2, 2, 256, 256, generator=gen)).cuda() * 0.3
    t = torch.tensor([0.9, 0.4], device="cuda")
    y = m(x, t)
    assert torch.isfinite(torch.view_as_real(y)).all()
    y_sw = m(x.flip(0), t.flip(0))
    assert torch.equal(y_sw.flip(0), y)


def _synth_Y(B, L):
    from storm_b200 import _native
    y = torch.stack([O.synth_utterance(b, L) for b in range(B)]).cuda()
    sp = _native.make_spec_config()
    norm = _native.wave_norm(y)
    T_pad = _native.padded_frames(_native.num_frames(L, sp))
    return y, norm, _native.stft(y, norm, sp, T_pad).unsqueeze(1)


def test_baseline_shape_single_nfe_vs_oracle():
    """BASELINE configs[1] shape (32640 samples -> 256x256, full-width ncsnpp): one network evaluation and one
    whole PC step (2 NFE, Philox noise replayed into the oracle) against the CPU oracle."""
    from storm_b200 import _native, ncsnpp_spec
    m = _model("ncsnpp")
    sd = ncsnpp_spec.random_state_dict(m.cfg, seed=0, init_scale=1.0, bias_std=0.05)
    m.load_state_dict(sd)
    _, _, Y = _synth_Y(1, 32640)
    assert tuple(Y.shape) == (1, 1, 256, 256)
    t = torch.tensor([0.7], device="cuda")
    x = Y + 0.1 * torch.view_as_complex(torch.randn(1, 1, 256, 256, 2, device="cuda", generator=torch.Generator(device="cuda").manual_seed(3)))
    d = m(torch.cat([x, Y], dim=1), t)
    torch.set_num_threads(min(32, len(__import__("os").sched_getaffinity(0))))
    do = O.ncsnpp_forward(sd, O.NetCfg(), torch.cat([x, Y], dim=1).cpu(), t.cpu())
    e1 = rel_l2(d.cpu(), do)
    print(f"256x256 single NFE rel-L2 vs oracle = {e1:.3e}")
    assert e1 < TOL
    sc = _native.make_sampler_config("reverse_diffusion", "ald", 1, 1, 0.5, 0.03)
    X, nfe = m.engine().pc_sample(Y, sc, seed=11, utt_ids=[5])
    n = Y[0].numel()
    it = iter([_native.philox_normal(11, 5, dr, n).reshape(1, 1, 256, 256).cpu() for dr in range(3)])
    Xo, _ = O.pc_sample(O.make_score_fn(sd, O.NetCfg()), Y.cpu(), O.OUVE(), 1, noise=lambda z: next(it))
    e2 = rel_l2(X.cpu(), Xo)
    print(f"256x256 one PC step rel-L2 vs oracle = {e2:.3e}")
    assert nfe == 2 and e2 < TOL


def test_trajectory_vs_oracle_and_sampler_sweep():
    """N-step trajectories (config 5 sweep at reduced size): PC vs fixed-step probability-flow, several N, against the
    oracle with replayed Philox noise; the error must not grow with N beyond the tolerance."""
    from storm_b200 import _native, ncsnpp_spec
    m = _model("ncsnpp")
    sd = ncsnpp_spec.random_state_dict(m.cfg, seed=0, init_scale=1.0, bias_std=0.05)
    m.load_state_dict(sd)
    _, _, Y = _synth_Y(2, 8000)          # 63 frames -> padded 64: [2,1,256,64]
    n = Y[0].numel()
    score = O.make_score_fn(sd, O.NetCfg())
    for N, pf, corr in [(4, False, "ald"), (8, False, "ald"), (6, True, "none")]:
        sc = _native.make_sampler_config("reverse_diffusion", corr, N, 1, 0.5, 0.03, probability_flow=pf)
        X, nfe = m.engine().pc_sample(Y, sc, seed=77, utt_ids=[0, 1])
        nd = 1 + N * ((1 if corr != "none" else 0) + 1)
        draws = [torch.stack([_native.philox_normal(77, u, dr, n).reshape(Y.shape[1:]) for u in (0, 1)]).cpu() for dr in range(nd)]
        it = iter(draws)
        Xo, nfeo = O.pc_sample(score, Y.cpu(), O.OUVE(), N, corrector=corr, noise=lambda z: next(it), probability_flow_discretize=pf)
        err = rel_l2(X.cpu(), Xo)
        print(f"N={N} pf={pf} corrector={corr}: nfe={nfe} rel-L2 vs oracle = {err:.3e}")
        assert nfe == nfeo and err < TOL


@pytest.mark.parametrize("name,L,T_pad", [("ncsnpplarge", 128000, 1024), ("ncsnpp", 64000, 512)])
def test_large_shapes_single_nfe_vs_oracle(name, L, T_pad):
    """BASELINE configs[3] shape (8 s -> 256x1024, ncsnpplarge) and the literal-4 s row (64000 samples -> 501 frames ->
    256x512, ncsnpp: 2048 attention tokens): one network evaluation against the CPU oracle."""
    from storm_b200 import ncsnpp_spec
    m = _model(name)
    sd = ncsnpp_spec.random_state_dict(m.cfg, seed=1, init_scale=1.0, bias_std=0.05)
    m.load_state_dict(sd)
    _, _, Y = _synth_Y(1, L)
    assert tuple(Y.shape) == (1, 1, 256, T_pad)
    t = torch.tensor([0.55], device="cuda")
    x = Y + 0.2 * torch.view_as_complex(torch.randn(1, 1, 256, T_pad, 2, device="cuda", generator=torch.Generator(device="cuda").manual_seed(5)))
    d = m(torch.cat([x, Y], dim=1), t)
    torch.set_num_threads(min(32, len(__import__("os").sched_getaffinity(0))))
    do = O.ncsnpp_forward(sd, O.named_cfg(name), torch.cat([x, Y], dim=1).cpu(), t.cpu())
    err = rel_l2(d.cpu(), do)
    print(f"{name} 256x{T_pad} single NFE rel-L2 vs oracle = {err:.3e}")
    assert err < TOL


def test_headline_batch_of_8_is_batch_invariant_and_matches_oracle():
    """The benchmarked batch (B = 8, 256x256, full-width ncsnpp): every utterance of the batch equals its own B = 1 run
    bit for bit (tile selection, GroupNorm statistics and Philox noise must not depend on the batch), and one of them
    matches the CPU oracle over a whole PC step."""
    from storm_b200 import _native, ncsnpp_spec
    m = _model("ncsnpp")
    sd = ncsnpp_spec.random_state_dict(m.cfg, seed=0, init_scale=1.0, bias_std=0.05)
    m.load_state_dict(sd)
    _, _, Y = _synth_Y(8, 32640)
    assert tuple(Y.shape) == (8, 1, 256, 256)
    sc = _native.make_sampler_config("reverse_diffusion", "ald", 1, 1, 0.5, 0.03)
    ids = [10, 11, 12, 13, 14, 15, 16, 17]
    X8, nfe = m.engine().pc_sample(Y, sc, seed=21, utt_ids=ids)
    assert nfe == 2
    # the same call again: now captured into a CUDA graph (second call) and replayed (third) - full-width network, every conv
    # schedule (CTA pairs, persistent tiles, one tile per CTA) inside the graph
    r0 = m.engine().graph_replays()
    for _ in range(2):
        Xg, _ = m.engine().pc_sample(Y, sc, seed=21, utt_ids=ids)
        assert torch.equal(Xg, X8), "graph replay differs from the eager trajectory"
    assert m.engine().graph_replays() == r0 + 2
    for b in (0, 3, 7):
        X1, _ = m.engine().pc_sample(Y[b:b + 1], sc, seed=21, utt_ids=[ids[b]])
        assert torch.equal(X1[0], X8[b]), f"utterance {b} depends on the batch"
    b = 3
    n = Y[0].numel()
    it = iter([_native.philox_normal(21, ids[b], dr, n).reshape(1, 1, 256, 256).cpu() for dr in range(3)])
    torch.set_num_threads(min(32, len(__import__("os").sched_getaffinity(0))))
    Xo, _ = O.pc_sample(O.make_score_fn(sd, O.NetCfg()), Y[b:b + 1].cpu(), O.OUVE(), 1, noise=lambda z: next(it))
    err = rel_l2(X8[b:b + 1].cpu(), Xo)
    print(f"B=8 headline shape, utterance {b}, one PC step rel-L2 vs oracle = {err:.3e}")
    assert err < TOL


def test_random_per_layer_weight_rescale_vs_oracle():
    """Every conv weight tensor rescaled by its own random 2^k, k in [-6, 6] (biases and GroupNorm untouched): the
    network function changes, and so do the magnitudes every tensor-core operand split sees - weights over 12 binades, the
    un-normalised residual stream over several.  The CUDA path must still match the oracle run on the same weights."""
    from storm_b200 import ncsnpp_spec
    m = _model("ncsnpp")
    sd = ncsnpp_spec.random_state_dict(m.cfg, seed=3, init_scale=1.0, bias_std=0.05)
    gen = torch.Generator().manual_seed(99)
    for k, v in sd.items():
        if k.endswith(".weight") and v.dim() == 4:      # the NIN maps feed a softmax: scaling q.k by 2^12 would test that, not the convs
            sd[k] = v * float(2.0 ** int(torch.randint(-6, 7, (1,), generator=gen)))
    m.load_state_dict(sd)
    g = torch.Generator().manual_seed(4)
    x = torch.complex(torch.randn(2, 2, 256, 64, generator=g), torch.randn(2, 2, 256, 64, generator=g)) * 0.4
    t = torch.tensor([0.9, 0.2])
    y = m(x.cuda(), t.cuda()).cpu()
    torch.set_num_threads(min(32, len(__import__("os").sched_getaffinity(0))))
    yo = O.ncsnpp_forward(sd, O.NetCfg(), x, t)
    err = rel_l2(y, yo)
    print(f"per-layer 2^[-6,6] weight rescale: rel-L2 vs oracle = {err:.3e}")
    assert torch.isfinite(torch.view_as_real(y)).all() and err < TOL


def test_ncsnpplarge_8s_shape_runs():
    """Config 4 shape: 128000 samples -> 1001 frames -> padded 1024; ncsnpplarge, one PC step, finite output and
    bit-exact independence from batch composition."""
    from storm_b200 import _native, ncsnpp_spec
    m = _model("ncsnpplarge")
    m.load_state_dict(ncsnpp_spec.random_state_dict(m.cfg, seed=1, init_scale=1.0))
    _, _, Y = _synth_Y(2, 128000)
    assert tuple(Y.shape) == (2, 1, 256, 1024)
    sc = _native.make_sampler_config("reverse_diffusion", "ald", 1, 1, 0.5, 0.03)
    X, nfe = m.engine().pc_sample(Y, sc, seed=3, utt_ids=[0, 1])
    assert nfe == 2 and torch.isfinite(torch.view_as_real(X)).all()
    X1, _ = m.engine().pc_sample(Y[1:2], sc, seed=3, utt_ids=[1])
    assert torch.equal(X1[0], X[1])


def test_storm_mode_vs_reference():
    """StochasticRegenerationModel.enhance (denoiser -> 6-channel score net around Y_denoised) vs the reference glue."""
    from storm_b200 import StochasticRegenerationModel
    g = load_golden("storm")
    model = StochasticRegenerationModel(backbone_denoiser="ncsnpp", backbone_score="ncsnpp", sde="ouve", nf=8)
    sd = {"denoiser_net." + k: v for k, v in sub_sd(g, "den/").items()}
    sd.update({"score_net." + k: v for k, v in sub_sd(g, "score/").items()})
    model.load_state_dict(sd)
    model.cuda()
    y = T(g["y"])
    sample, Y, T_orig, norm = model.enhance(y, N=int(g["N"]), return_stft=True, noise=_noise(g["noise"]))
    assert T_orig == y.shape[1] and rel_l2(Y.cpu(), g["Y"].squeeze()) < 1e-5
    assert rel_l2(sample.cpu(), g["sample"].squeeze()) < TIGHT
    den = model.enhance(y, denoiser_only=True, return_stft=True)[0]
    assert rel_l2(den.cpu(), g["Y_den"].squeeze()) < TIGHT
    # waveform output with Philox noise: finite, right length, on the CPU like the reference
    x_hat, nfe, rtf = model.enhance(y, N=2, timeit=True)
    assert x_hat.shape == (y.shape[1],) and x_hat.device.type == "cpu" and nfe == 2 and torch.isfinite(x_hat).all()


def test_storm_enhance_batch_equals_per_file_and_oracle():
    """Batched StoRM (one denoiser forward + one native sampler for the whole batch, ragged lengths inside a bucket)
    equals enhancing every file alone with the same seed / utterance ids bit for bit, and matches the CPU oracle."""
    from storm_b200 import StochasticRegenerationModel, _native, ncsnpp_spec
    model = StochasticRegenerationModel(backbone_denoiser="ncsnpp", backbone_score="ncsnpp", sde="ouve", nf=32, seed=5)
    sd_d = ncsnpp_spec.random_state_dict(model.denoiser_net.cfg, seed=11, init_scale=1.0, bias_std=0.05)
    sd_s = ncsnpp_spec.random_state_dict(model.score_net.cfg, seed=12, init_scale=1.0, bias_std=0.05)
    model.denoiser_net.load_state_dict(sd_d)
    model.score_net.load_state_dict(sd_s)
    model.cuda()
    lens = [8000, 7700, 8100, 16000]                         # three share the 64-frame bucket, one is alone in the 128-frame one
    ys = [O.synth_utterance(40 + i, L) for i, L in enumerate(lens)]
    outs, nfe = model.enhance_batch(ys, N=3, seed=77, utt_ids=[100, 101, 102, 103], return_nfe=True)
    assert nfe == 3 and [o.numel() for o in outs] == lens and all(o.device.type == "cpu" for o in outs)
    for i, y in enumerate(ys):
        one = model.enhance(y.reshape(1, -1), N=3, seed=77, utt_ids=[100 + i])
        assert torch.equal(one, outs[i]), f"utterance {i}: batched result differs from the per-file result"
    # oracle for the first utterance, Philox noise replayed (1 prior draw + N predictor draws)
    y0 = ys[0].reshape(1, -1)
    sp = _native.make_spec_config()
    T_pad = _native.padded_frames(_native.num_frames(lens[0], sp))
    n = 256 * T_pad
    it = iter([_native.philox_normal(77, 100, d, n).reshape(1, 1, 256, T_pad).cpu() for d in range(4)])
    xo, nfeo = O.enhance_storm(sd_d, sd_s, O.named_cfg("ncsnpp", nf=32, input_channels=2, discriminative=True),
                               O.named_cfg("ncsnpp", nf=32, input_channels=6), y0, N=3, noise=lambda z: next(it))
    err = rel_l2(outs[0], xo)
    print(f"StoRM batch, utterance 0 waveform rel-L2 vs oracle = {err:.3e}")
    assert nfeo == 3 and err < TOL


def test_discriminative_model_vs_oracle(tmp_path):
    """DiscriminativeModel (sgmse/model.py:320-370, --mode denoiser-only): ScoreModel-shaped checkpoint (``dnn.*`` keys),
    one network call with t = 1, batched."""
    from storm_b200 import DiscriminativeModel, ncsnpp_spec
    src = DiscriminativeModel(backbone="ncsnpp", sde="ouve", nf=32, discriminative=True, init_scale=1.0)
    assert src.dnn.cfg.in_ch == 2
    sd = ncsnpp_spec.random_state_dict(src.dnn.cfg, seed=21, init_scale=1.0, bias_std=0.05)
    src.dnn.load_state_dict(sd)
    ckpt = dict(state_dict=src.state_dict(), hyper_parameters=dict(backbone="ncsnpp", sde="ouve", nf=32, discriminative=True, t_eps=0.03))
    path = tmp_path / "disc.ckpt"
    torch.save(ckpt, path)
    m = DiscriminativeModel.load_from_checkpoint(str(path), base_dir="", batch_size=1, num_workers=0, kwargs=dict(gpu=False))
    m.eval(no_ema=False).cuda()
    ys = torch.stack([O.synth_utterance(60 + i, 8000) for i in range(3)])
    out = m.enhance_batch(ys)
    assert out.shape == ys.shape and out.device == ys.device
    one = m.enhance(ys[1:2])
    assert one.device.type == "cpu" and torch.equal(one, out[1])
    sc = O.SpecCfg()
    y = ys[1:2]
    nf = y.abs().max()
    Y = O.pad_spec(O.spec_fwd(O.stft(y / nf, sc), sc).unsqueeze(0))
    Xo = O.ncsnpp_forward(sd, O.named_cfg("ncsnpp", nf=32, input_channels=2, discriminative=True), Y, None)
    xo = O.istft(O.spec_back(Xo.squeeze(), sc), sc, 8000) * nf
    err = rel_l2(out[1], xo.squeeze())
    print(f"DiscriminativeModel waveform rel-L2 vs oracle = {err:.3e}")
    assert err < TOL
    with pytest.raises(RuntimeError, match="real input channels"):
        DiscriminativeModel(backbone="ncsnpp", sde="ouve", nf=32).cuda().enhance_batch(ys)     # stock 4-channel backbone: same mismatch as model.py:326


@pytest.mark.parametrize("frozen_has_shadow", [True, False])
def test_reference_checkpoint_with_ema_runs_on_the_engine(tmp_path, frozen_has_shadow):
    """A Lightning checkpoint in the reference's layout (``dnn.*`` state_dict in module order, ``ema.shadow_params`` in
    ``parameters()`` order - tests/golden/param_order.json - with or without a shadow for the frozen Fourier W, the two
    torch_ema conventions) -> ScoreModel.load_from_checkpoint -> eval() swaps the shadow weights into the native engine:
    one PC step against the oracle run on the SHADOW weights; eval(no_ema=True) against the raw ones."""
    import json
    import os
    from conftest import GOLDEN
    from storm_b200 import ScoreModel, _native, ncsnpp_spec
    cfg = ncsnpp_spec.NCSNppConfig(nf=32)
    raw = ncsnpp_spec.random_state_dict(cfg, seed=31, init_scale=1.0, bias_std=0.05)
    ema = ncsnpp_spec.random_state_dict(cfg, seed=32, init_scale=1.0, bias_std=0.05)
    ema["all_modules.0.W"] = raw["all_modules.0.W"].clone()             # frozen: never updated, identical in both sets
    order = json.load(open(os.path.join(GOLDEN, "param_order.json")))["ncsnpp_4ch"]
    assert [n for n, _, _ in order] == list(raw.keys())
    shadow = [ema[n].clone() for n, _, rg in order if rg or frozen_has_shadow]
    ckpt = dict(state_dict={"dnn." + k: v for k, v in raw.items()},
                hyper_parameters=dict(backbone="ncsnpp", sde="ouve", nf=32, t_eps=0.03, theta=1.5, sigma_min=0.05, sigma_max=0.5,
                                      data_module_cls=None),
                ema=dict(decay=0.999, num_updates=1000, shadow_params=shadow, collected_params=None))
    path = tmp_path / "score.ckpt"
    torch.save(ckpt, path)
    m = ScoreModel.load_from_checkpoint(str(path), base_dir="", batch_size=1, num_workers=0, kwargs=dict(gpu=False))
    m.eval(no_ema=False)
    m.cuda()
    _, _, Y = _synth_Y(1, 8000)
    sc = _native.make_sampler_config("reverse_diffusion", "ald", 1, 1, 0.5, 0.03)
    n = Y[0].numel()
    draws = [_native.philox_normal(13, 2, d, n).reshape(Y.shape).cpu() for d in range(3)]
    ocfg = O.named_cfg("ncsnpp", nf=32)
    for sd_o, no_ema in ((ema, False), (raw, True)):
        m.train(True)
        m.eval(no_ema=no_ema)
        X, nfe = m.dnn.engine().pc_sample(Y, sc, seed=13, utt_ids=[2])
        it = iter(draws)
        Xo, _ = O.pc_sample(O.make_score_fn(sd_o, ocfg), Y.cpu(), O.OUVE(), 1, noise=lambda z: next(it))
        err = rel_l2(X.cpu(), Xo)
        print(f"checkpoint (frozen W shadowed: {frozen_has_shadow}), no_ema={no_ema}: one PC step rel-L2 vs oracle = {err:.3e}")
        assert nfe == 2 and err < TOL
    other = rel_l2(X.cpu(), O.pc_sample(O.make_score_fn(ema, ocfg), Y.cpu(), O.OUVE(), 1, noise=lambda z, it=iter(draws): next(it))[0])
    assert other > 10 * TOL        # the two weight sets really differ: a mis-assigned shadow list could not pass both checks


@pytest.mark.parametrize("pred,corr", [("reverse_diffusion", "ald"), ("euler_maruyama", "langevin"), ("reverse_diffusion", "none")])
def test_graph_replay_is_bit_identical_to_the_eager_loop(pred, corr):
    """the N-step loop captured into a CUDA graph (second call of a shape) and replayed (later calls) gives
    the eager loop's result bit for bit, with seed and utterance ids - read from device memory by the replayed noise
    kernels - changing between calls."""
    from storm_b200 import _native, ncsnpp_spec
    m = _model("ncsnpp", nf=32)
    m.load_state_dict(ncsnpp_spec.random_state_dict(m.cfg, seed=8, init_scale=1.0, bias_std=0.05))
    eng = m.engine()
    _, _, Y = _synth_Y(3, 8000)
    sc = _native.make_sampler_config(pred, corr, 4, 1, 0.5, 0.03)
    calls = [(1, [0, 1, 2]), (2, [7, 8, 9]), (3, [4, 4, 6]), (2, [7, 8, 9])]
    eng.set_graph(False)
    eager = [eng.pc_sample(Y * (1.0 + 0.1 * i), sc, seed=sd, utt_ids=ids)[0].clone() for i, (sd, ids) in enumerate(calls)]
    eng.set_graph(True)
    r0 = eng.graph_replays()
    for i, (sd, ids) in enumerate(calls):
        X, nfe = eng.pc_sample(Y * (1.0 + 0.1 * i), sc, seed=sd, utt_ids=ids)
        assert nfe == 4 * (2 if corr != "none" else 1)
        assert torch.equal(X, eager[i]), f"call {i}: graph path differs from the eager path"
    assert eng.graph_replays() - r0 == len(calls) - 1          # first call eager (sizes the workspace), then capture + replays
    # a new batch size gets its own graph; the old one keeps working
    X1, _ = eng.pc_sample(Y[:1], sc, seed=2, utt_ids=[7])
    X1b, _ = eng.pc_sample(Y[:1], sc, seed=2, utt_ids=[7])
    assert torch.equal(X1, X1b)
    if corr != "langevin":                                      # langevin couples the batch (correctors.py:53-55)
        assert torch.equal(X1[0], eager[1][0] if False else eng.pc_sample(Y, sc, seed=2, utt_ids=[7, 8, 9])[0][0])
    # the whole enhance() call (STFT -> replayed sampler -> iSTFT)
    y = torch.stack([O.synth_utterance(b, 8000) for b in range(3)]).cuda()
    sp = _native.make_spec_config()
    a, _ = eng.enhance(y, sc, sp, seed=5, utt_ids=[1, 2, 3])
    b, _ = eng.enhance(y, sc, sp, seed=5, utt_ids=[1, 2, 3])
    eng.set_graph(False)
    c, _ = eng.enhance(y, sc, sp, seed=5, utt_ids=[1, 2, 3])
    eng.set_graph(True)
    assert torch.equal(a, c) and torch.equal(b, c)


def test_engine_follows_the_weights_to_another_device_or_complains():
    """Operands on a device other than the engine's are reported, not launched on (device handling)."""
    from storm_b200 import ncsnpp_spec
    m = _model("ncsnpp", nf=32)
    m.load_state_dict(ncsnpp_spec.random_state_dict(m.cfg, seed=2, init_scale=1.0))
    x = torch.complex(torch.randn(1, 2, 64, 64), torch.randn(1, 2, 64, 64)).cuda()
    t = torch.tensor([0.5]).cuda()
    y0 = m(x, t)
    assert m.engine().device == x.device
    if torch.cuda.device_count() > 1:
        m.cuda(1)
        y1 = m(x.to("cuda:1"), t.to("cuda:1"))                      # current device is still 0
        assert m.engine().device == torch.device("cuda", 1) and torch.equal(y1.cpu(), y0.cpu())
        with pytest.raises(RuntimeError, match="created on"):
            m(x, t)


@pytest.mark.parametrize("name", ["ncsnpp12M", "ncsnpp6M"])
def test_small_backbones_vs_oracle(name):
    """nf = 96 variants (ncsnpp.py:479-513): channel counts 96/192 exercise zero-padded 128-row weight blocks."""
    from storm_b200 import ncsnpp_spec
    m = _model(name)
    sd = ncsnpp_spec.random_state_dict(m.cfg, seed=4, init_scale=1.0, bias_std=0.05)
    m.load_state_dict(sd)
    gen = torch.Generator().manual_seed(5)
    x = torch.complex(torch.randn(2, 2, 64, 64, generator=gen), torch.randn(2, 2, 64, 64, generator=gen)) * 0.4
    t = torch.tensor([0.9, 0.05])
    y = m(x.cuda(), t.cuda()).cpu()
    yo = O.ncsnpp_forward(sd, O.named_cfg(name), x, t)
    err = rel_l2(y, yo)
    print(f"{name}: rel-L2 vs oracle = {err:.3e}")
    assert err < TOL


def test_error_behaviour_matches_reference_conventions():
    from storm_b200 import BackboneRegistry, ScoreModel, _native
    m = _model("ncsnpp", nf=8, init_scale=1.0)
    with pytest.raises(ValueError):           # T not a multiple of 2^(levels-1): the reference's skip-cat would fail too
        m(torch.zeros(1, 2, 16, 20, dtype=torch.complex64, device="cuda"), torch.ones(1, device="cuda"))
    with pytest.raises(ValueError):           # wrong channel count
        m(torch.zeros(1, 3, 16, 16, dtype=torch.complex64, device="cuda"), torch.ones(1, device="cuda"))
    with pytest.raises(ValueError):           # missing time conditioning on a score network
        m(torch.zeros(1, 2, 16, 16, dtype=torch.complex64, device="cuda"))
    with pytest.raises(RuntimeError):         # host tensors are refused, no silent CPU path
        m(torch.zeros(1, 2, 16, 16, dtype=torch.complex64), torch.ones(1))
    sc = _native.make_sampler_config("reverse_diffusion", "ald", 2)
    with pytest.raises(ValueError):           # injected noise with the wrong number of draws
        m.engine().pc_sample(torch.zeros(1, 1, 16, 16, dtype=torch.complex64, device="cuda"), sc,
                             noise=torch.zeros(2, 1, 1, 16, 16, dtype=torch.complex64, device="cuda"))
    with pytest.raises(ValueError):
        _native.make_sampler_config("reverse_diffusion", "ald", 0)._fields_ and m.engine().pc_sample(
            torch.zeros(1, 1, 16, 16, dtype=torch.complex64, device="cuda"), _native.make_sampler_config("reverse_diffusion", "ald", 0))
    with pytest.raises(ValueError):           # unknown backbone name through the model surface
        ScoreModel(backbone="nope")


def test_ragged_batch_equals_per_file_enhancement():
    """Utterances of different lengths in one padded-frame bucket share a batch (storm_enhance_ragged) and still get the
    frames, reflect padding, pad_spec amount, normalisation and istft length they get when enhanced alone."""
    from storm_b200 import ScoreModel, _native
    model = ScoreModel(backbone="ncsnpp", sde="ouve", nf=8, init_scale=1.0).cuda()
    sp = model.data_module.spec_config()
    lens = [5000, 4000, 7000, 9000, 8100]          # T_pad 64, 64, 64, 128, 64
    ys = [O.synth_utterance(b, L) for b, L in enumerate(lens)]
    # front end: ragged STFT rows == per-utterance STFT (index work and values exact)
    idx = [0, 1, 2, 4]
    Lmax = max(lens[i] for i in idx)
    rows = torch.zeros(len(idx), Lmax)
    for r, i in enumerate(idx):
        rows[r, :lens[i]] = ys[i]
    rows = rows.cuda()
    ln = [lens[i] for i in idx]
    norm = _native.wave_norm(rows, lengths=ln)
    Y = _native.stft(rows, norm, sp, 64, lengths=ln)
    for r, i in enumerate(idx):
        yi = ys[i].reshape(1, -1).cuda()
        ni = _native.wave_norm(yi)
        assert float(ni[0]) == float(norm[r]) == float(ys[i].abs().max())
        Yi = _native.stft(yi, ni, sp, 64)
        assert torch.equal(Yi[0], Y[r])
        frames = 1 + lens[i] // 128
        assert torch.count_nonzero(Y[r, :, frames:]) == 0
    with pytest.raises(ValueError):
        _native.stft(rows, norm, sp, 64, lengths=[5000, 4000, 7000, 100])      # shorter than the reflect padding
    with pytest.raises(ValueError):
        _native.stft(rows, norm, sp, 64, lengths=[5000, 4000, 7000, 9000])     # longer than the row
    # whole path: list API buckets by frames, results equal the one-at-a-time results
    outs, nfe = model.enhance_batch(ys, N=2, seed=9, utt_ids=list(range(len(ys))), return_nfe=True, max_batch=3)
    assert nfe == 4 and [o.numel() for o in outs] == lens
    for i, y in enumerate(ys):
        alone = model.enhance_batch(y.reshape(1, -1), N=2, seed=9, utt_ids=[i])[0]
        assert outs[i].shape == alone.shape
        assert rel_l2(outs[i].cpu(), alone.cpu()) < 1e-5
    # rows shorter than the stride are zero beyond their own length in the raw ragged output
    out, _ = model.dnn.engine().enhance(rows, model._sampler_cfg("reverse_diffusion", "ald", 2, 1, 0.5), sp, seed=9, utt_ids=idx, lengths=ln)
    for r, i in enumerate(idx):
        assert torch.count_nonzero(out[r, lens[i]:]) == 0


@pytest.mark.parametrize("tag", ["a", "b", "c", "d"])
def test_si_sdr_vs_reference(tag):
    from storm_b200.util.other import si_sdr, si_sdr_torch
    g = load_golden("metrics")
    s, sh = T(g[f"{tag}/s"]), T(g[f"{tag}/s_hat"])
    assert abs(si_sdr(s.numpy(), sh.numpy()) - float(g[f"{tag}/si_sdr"])) < 2e-3      # dB; inputs are fp32 here, fp64 there
    assert abs(float(si_sdr_torch(s.cuda(), sh.cuda())) - float(g[f"{tag}/si_sdr_torch"])) < 2e-3


def test_si_sdr_batch_ragged_and_truncation():
    from storm_b200 import _native
    from storm_b200.util.other import si_sdr_torch
    g = load_golden("metrics")
    assert abs(float(si_sdr_torch(T(g["e/s"]).cuda(), T(g["e/s_hat"]).cuda())) - float(g["e/si_sdr_torch"])) < 2e-3
    tags = ["a", "c", "d"]
    L = max(g[f"{t}/s"].shape[0] for t in tags)
    S, H = torch.zeros(3, L), torch.ones(3, L)          # garbage beyond each row's length must be ignored
    for r, t in enumerate(tags):
        n = g[f"{t}/s"].shape[0]
        S[r, :n], H[r, :n] = T(g[f"{t}/s"]), T(g[f"{t}/s_hat"])
    out = _native.si_sdr(S.cuda(), H.cuda(), lengths=[g[f"{t}/s"].shape[0] for t in tags]).cpu()
    for r, t in enumerate(tags):
        assert abs(float(out[r]) - float(g[f"{t}/si_sdr"])) < 2e-3


def test_cli_enhances_a_directory_of_ragged_files(tmp_path):
    """enhancement.py end to end (reference flags, enhancement.py:28-36) on wav files of different lengths: every file is
    written with its own length and equals what ScoreModel.enhance_batch gives for that utterance alone."""
    import os
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import enhancement
    from storm_b200 import ScoreModel
    from storm_b200.util import wavio
    src, dst = tmp_path / "noisy", tmp_path / "out"
    src.mkdir()
    lens = [6000, 5000, 9000, 6100]
    for i, L in enumerate(lens):
        wavio.save_wav(str(src / f"f{i}.wav"), O.synth_utterance(i, L).reshape(1, -1), 16000)
    enhancement.main(["--test_dir", str(src), "--enhanced_dir", str(dst), "--random-init", "ncsnpp6M", "--N", "1",
                      "--batch-size", "3", "--seed", "11"])
    model = ScoreModel(backbone="ncsnpp6M", sde="ouve", init_scale=1.0, seed=11).cuda()
    for i, L in enumerate(lens):
        y, sr = wavio.load_wav(str(dst / f"f{i}.wav"))
        assert sr == 16000 and y.shape == (1, L) and torch.isfinite(y).all()
        x, _ = wavio.load_wav(str(src / f"f{i}.wav"))
        alone = model.enhance_batch(x, N=1, seed=11, utt_ids=[i])
        assert rel_l2(y, alone.cpu()) < 1e-4


def test_headline_config_vs_reference():
    """BASELINE configs[1] for one utterance against the UNMODIFIED reference (fixture headline_256x256_N30, generated by
    oracle/gen_golden.py): 32 640 samples -> 256x256, full-width ncsnpp, N=30 reverse_diffusion + ald (60 network
    evaluations) with the same injected noise; weights, waveform and noise are regenerated from their seeds."""
    from storm_b200 import _native, ncsnpp_spec
    g = load_golden("headline_256x256_N30")
    m = _model("ncsnpp")
    m.load_state_dict(ncsnpp_spec.random_state_dict(m.cfg, seed=int(g["weight_seed"]), init_scale=float(g["init_scale"]),
                                                    bias_std=float(g["bias_std"])))
    sp = _native.make_spec_config()
    L, N = int(g["L"]), int(g["N"])
    y = O.synth_utterance(0, L).reshape(1, -1).cuda()
    norm = _native.wave_norm(y)
    T_pad = _native.padded_frames(_native.num_frames(L, sp))
    assert T_pad == 256                                                  # index work: exact
    Y = _native.stft(y, norm, sp, T_pad).unsqueeze(1)
    zs = torch.stack(O.seeded_complex_noise(int(g["noise_seed"]), 1 + 2 * N, (1, 1, 256, 256))).cuda()
    sc = _native.make_sampler_config("reverse_diffusion", "ald", N, 1, 0.5, 0.03)
    X, nfe = m.engine().pc_sample(Y, sc, noise=zs)
    assert nfe == int(g["nfe"]) == 60
    err = rel_l2(X.cpu(), g["sample"])
    x_hat = m.engine().istft(X.squeeze(1), norm, L, sp).squeeze().cpu()
    err_w = rel_l2(x_hat, g["x_hat"])
    print(f"headline config vs reference: enhanced spectrogram rel-L2 = {err:.3e}, waveform rel-L2 = {err_w:.3e}")
    assert err < TOL and err_w < TOL
    assert err < TIGHT                                                   # what the bf16x3 path is expected to reach
