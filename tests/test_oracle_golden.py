"""Pin the CPU oracle (oracle/storm_oracle.py) against fixtures produced by the unmodified reference
(oracle/gen_golden.py).  CPU only."""
import numpy as np
import pytest
import torch

from conftest import load_golden, rel_l2, sub_sd
from oracle import storm_oracle as O
from storm_b200 import ncsnpp_spec

T = torch.from_numpy
TOL = 2e-5


def test_fir_closed_forms():
    g = load_golden("fir")
    for x, up, down in [("x", "up", "down"), ("x1", "up1", "down1")]:
        assert rel_l2(O.fir_upsample(T(g[x])), g[up]) < 1e-6
        assert rel_l2(O.fir_downsample(T(g[x])), g[down]) < 1e-6
        assert O.fir_upsample(T(g[x])).shape == g[up].shape and O.fir_downsample(T(g[x])).shape == g[down].shape


@pytest.mark.parametrize("name,up,down", [("plain", 0, 0), ("chg", 0, 0), ("down", 0, 1), ("up", 1, 0), ("cat", 0, 0)])
def test_resblock(name, up, down):
    g = load_golden("layers")
    sd = sub_sd(g, f"{name}/sd/")
    y = O.resblock(T(g[f"{name}/x"]), T(g[f"{name}/temb"]), sd, "blk", up=bool(up), down=bool(down))
    assert rel_l2(y, g[f"{name}/y"]) < TOL


def test_attnblock():
    g = load_golden("layers")
    y = O.attnblock(T(g["attn/x"]), sub_sd(g, "attn/sd/"), "blk")
    assert rel_l2(y, g["attn/y"]) < TOL


@pytest.mark.parametrize("fix,cfg", [
    ("ncsnpp_tiny", O.NetCfg(nf=8)),
    ("ncsnpp_tiny_6ch", O.NetCfg(nf=8, input_channels=6)),
    ("ncsnpp_tiny_disc", O.NetCfg(nf=8, discriminative=True)),
    ("ncsnpplarge_tiny", O.named_cfg("ncsnpplarge", nf=8)),
])
def test_network_tiny(fix, cfg):
    g = load_golden(fix)
    sd = sub_sd(g, "sd/")
    t = T(g["t"]) if "t" in g else None
    y = O.ncsnpp_forward(sd, cfg, T(g["x"]), t)
    assert y.shape == g["y"].shape
    assert rel_l2(y, g["y"]) < 5e-5
    assert O.count_modules(cfg) == 1 + max(int(k.split(".")[1]) for k in sd if k.startswith("all_modules."))


@pytest.mark.parametrize("fix,name", [("ncsnpp_full_64x64", "ncsnpp"), ("ncsnpplarge_full_256x64", "ncsnpplarge")])
def test_network_full_width_seeded(fix, name):
    g = load_golden(fix)
    ocfg = O.named_cfg(name)
    pcfg = ncsnpp_spec.NCSNppConfig(nf=ocfg.nf, ch_mult=ocfg.ch_mult, num_res_blocks=ocfg.num_res_blocks,
                                    attn_resolutions=ocfg.attn_resolutions)
    sd = ncsnpp_spec.random_state_dict(pcfg, seed=int(g["seed"]), init_scale=float(g["init_scale"]), bias_std=float(g["bias_std"]))
    y = O.ncsnpp_forward(sd, ocfg, T(g["x"]), T(g["t"]))
    assert rel_l2(y, g["y"]) < 5e-5
    # non-degenerate: the output really depends on the input
    assert float(torch.as_tensor(g["y"]).abs().std()) > 1e-3


def _queue(noise):
    zs = [T(np.ascontiguousarray(z)) for z in noise]
    it = iter(zs)
    return lambda x: next(it)


@pytest.mark.parametrize("tag,pred,corr", [("rd_ald", "reverse_diffusion", "ald"), ("rd_ald2", "reverse_diffusion", "ald"),
                                           ("em_langevin", "euler_maruyama", "langevin"),
                                           ("rd_none", "reverse_diffusion", "none"), ("none_ald", "none", "ald")])
def test_pc_sampler(tag, pred, corr):
    g = load_golden("sampler")
    sd = sub_sd(g, "sd/")
    score = O.make_score_fn(sd, O.NetCfg(nf=8))
    x, nfe = O.pc_sample(score, T(g["y"]), O.OUVE(), int(g[f"{tag}/N"]), pred, corr, int(g[f"{tag}/csteps"]),
                         float(g[f"{tag}/snr"]), 0.03, _queue(g[f"{tag}/noise"]))
    assert nfe == int(g[f"{tag}/nfe"])
    assert rel_l2(x, g[f"{tag}/x"]) < 5e-5


def test_sde_scalars():
    g = load_golden("sampler")
    s = O.OUVE()
    t = T(g["sde/t"])
    np.testing.assert_allclose(s.std(t).numpy(), g["sde/std"], rtol=1e-6)
    np.testing.assert_allclose(s.diffusion(t).numpy(), g["sde/diffusion"], rtol=1e-6)


def test_ode_sampler():
    g = load_golden("sampler")
    score = O.make_score_fn(sub_sd(g, "sd/"), O.NetCfg(nf=8))
    x, nfe = O.ode_sample(score, T(g["y"]), O.OUVE(), int(g["ode/N"]), 0.03, noise=_queue(g["ode/noise"]))
    assert nfe == int(g["ode/nfe"])
    assert rel_l2(x, g["ode/x"]) < 1e-4


@pytest.mark.parametrize("tag", ["a", "b", "c"])
def test_stft_chain(tag):
    g = load_golden("stft")
    sc = O.SpecCfg()
    sig = T(g[f"{tag}/sig"])
    S = O.stft(sig, sc)
    assert S.shape == g[f"{tag}/stft"].shape                       # frame count: bit-exact index work
    assert S.shape[-1] == 1 + sig.shape[-1] // sc.hop_length and S.shape[-2] == 256
    assert rel_l2(S, g[f"{tag}/stft"]) < 1e-6
    Sf = O.spec_fwd(S, sc)
    assert rel_l2(Sf, g[f"{tag}/fwd"]) < 1e-6
    assert O.pad_spec(Sf.unsqueeze(0)).shape[-1] == int(g[f"{tag}/padded_T"])
    back = O.istft(O.spec_back(Sf, sc), sc, sig.shape[-1])
    assert back.shape == g[f"{tag}/back"].shape
    assert rel_l2(back, g[f"{tag}/back"]) < 1e-6


def test_stft_other_transform():
    g = load_golden("stft")
    sc = O.SpecCfg(spec_factor=0.33, spec_abs_exponent=1.0, window="sqrthann")
    S = O.stft(T(g["d/sig"]), sc)
    assert rel_l2(S, g["d/stft"]) < 1e-6
    assert rel_l2(O.spec_fwd(S, sc), g["d/fwd"]) < 1e-6
    assert rel_l2(O.istft(O.spec_back(O.spec_fwd(S, sc), sc), sc, 3000), g["d/back"]) < 1e-6


def test_enhance_end_to_end():
    g = load_golden("enhance")
    sd = sub_sd(g, "sd/")
    x_hat, nfe = O.enhance(sd, O.NetCfg(nf=8), T(g["y"]), N=int(g["N"]), noise=_queue(g["noise"]))
    assert nfe == int(g["nfe"])
    assert x_hat.shape == g["x_hat"].shape
    assert rel_l2(x_hat, g["x_hat"]) < 5e-5
    sample, Y, T_orig, norm = O.enhance(sd, O.NetCfg(nf=8), T(g["y"]), N=int(g["N"]), noise=_queue(g["noise"]), return_stft=True)
    assert T_orig == g["y"].shape[1] and rel_l2(Y, g["Y"].squeeze()) < 1e-6


def test_param_table_matches_reference_keys():
    g = load_golden("ncsnpp_tiny")
    keys = [k[3:] for k in g if k.startswith("sd/")]
    tab = ncsnpp_spec.param_table(ncsnpp_spec.NCSNppConfig(nf=8))
    assert [k for k, _ in tab] == keys
    assert all(tuple(g["sd/" + k].shape) == s for k, s in tab)
    gl = load_golden("ncsnpplarge_tiny")
    tabl = ncsnpp_spec.param_table(ncsnpp_spec.NCSNppConfig(nf=8, ch_mult=(1, 1, 2, 2, 2, 2, 2), num_res_blocks=2, attn_resolutions=(16,)))
    assert [k for k, _ in tabl] == [k[3:] for k in gl if k.startswith("sd/")]
    assert ncsnpp_spec.param_count(ncsnpp_spec.NCSNppConfig()) == 27756314


def test_storm_enhance_end_to_end():
    g = load_golden("storm")
    x_hat, nfe = O.enhance_storm(sub_sd(g, "den/"), sub_sd(g, "score/"), O.NetCfg(nf=8, discriminative=True),
                                 O.NetCfg(nf=8, input_channels=6), T(g["y"]), N=int(g["N"]), noise=_queue(g["noise"]))
    assert nfe == int(g["nfe"]) == 3
    assert rel_l2(x_hat, g["x_hat"]) < 5e-5
    sample, Y, Y_den, _, _ = O.enhance_storm(sub_sd(g, "den/"), sub_sd(g, "score/"), O.NetCfg(nf=8, discriminative=True),
                                             O.NetCfg(nf=8, input_channels=6), T(g["y"]), N=int(g["N"]), noise=_queue(g["noise"]),
                                             return_stft=True)
    assert rel_l2(Y_den, g["Y_den"].squeeze()) < 5e-5 and rel_l2(sample, g["sample"].squeeze()) < 5e-5


@pytest.mark.parametrize("tag", ["a", "b", "c", "d"])
def test_si_sdr(tag):
    g = load_golden("metrics")
    s, sh = g[f"{tag}/s"], g[f"{tag}/s_hat"]
    assert abs(O.si_sdr(s, sh) - float(g[f"{tag}/si_sdr"])) < 1e-9
    got = O.si_sdr_torch(torch.from_numpy(s), torch.from_numpy(sh))
    assert abs(float(got) - float(g[f"{tag}/si_sdr_torch"])) < 1e-4


def test_si_sdr_torch_truncates_to_common_length():
    g = load_golden("metrics")
    got = O.si_sdr_torch(torch.from_numpy(g["e/s"]), torch.from_numpy(g["e/s_hat"]))
    assert abs(float(got) - float(g["e/si_sdr_torch"])) < 1e-4
