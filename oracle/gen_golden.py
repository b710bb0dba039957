"""Generate the committed golden fixtures under tests/golden/ by running the UNMODIFIED reference.

TEST INFRASTRUCTURE ONLY.  Runs only in the build container (needs /root/reference); the fixtures it
writes are what travels to the GPU box.  Usage:  python oracle/gen_golden.py [--only NAME]

Each fixture stores the seeded inputs, (small) weights where the model is tiny, and the reference's
outputs.  The reference ships no golden vectors of its own, so these are the pins.
Noise is injected by temporarily replacing ``torch.randn_like`` with a queue of pre-generated tensors,
which the reference's sampler code reaches through the ``torch`` module attribute
(sdes.py:237, predictors.py:48,64, correctors.py:52,86).
"""
import argparse
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ref_shim  # noqa: E402
from storm_b200 import ncsnpp_spec  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def c2np(x):
    return x.detach().cpu().numpy()


def save(name, **arrs):
    path = os.path.join(OUT, name + ".npz")
    np.savez_compressed(path, **{k: (c2np(v) if torch.is_tensor(v) else np.asarray(v)) for k, v in arrs.items()})
    print(f"wrote {path}  ({os.path.getsize(path) / 1024:.0f} KiB)")


def sd_arrays(sd, prefix="sd/"):
    return {prefix + k: v for k, v in sd.items()}


def randomize_(module, gen, scale=0.3):
    """Fill every parameter of a reference module with non-degenerate seeded values."""
    with torch.no_grad():
        for n, p in module.named_parameters():
            if n.endswith("GroupNorm_0.weight") or n.endswith("GroupNorm_1.weight") or (p.dim() == 1 and "weight" in n and "Norm" in n):
                p.copy_(1.0 + 0.2 * torch.randn(p.shape, generator=gen))
            else:
                p.copy_(scale * torch.randn(p.shape, generator=gen))


class NoiseQueue:
    def __init__(self, tensors):
        self.q = list(tensors)
        self.i = 0

    def __call__(self, x, **kw):
        z = self.q[self.i]
        self.i += 1
        assert z.shape == x.shape and z.dtype == x.dtype
        return z


class patched_randn_like:
    def __init__(self, q):
        self.q = q

    def __enter__(self):
        self.orig = torch.randn_like
        torch.randn_like = self.q

    def __exit__(self, *a):
        torch.randn_like = self.orig


def cplx(gen, *shape, scale=1.0):
    return torch.complex(torch.randn(*shape, generator=gen), torch.randn(*shape, generator=gen)) * scale


# --------------------------------------------------------------------------------------------------
def gen_fir(ns):
    g = torch.Generator().manual_seed(1)
    x = torch.randn(2, 3, 6, 10, generator=g)
    x1 = torch.randn(1, 2, 2, 2, generator=g)
    save("fir", x=x, up=ns.up_or_down_sampling.upsample_2d(x, [1, 3, 3, 1], factor=2),
         down=ns.up_or_down_sampling.downsample_2d(x, [1, 3, 3, 1], factor=2),
         x1=x1, up1=ns.up_or_down_sampling.upsample_2d(x1, [1, 3, 3, 1], factor=2),
         down1=ns.up_or_down_sampling.downsample_2d(x1, [1, 3, 3, 1], factor=2))


def gen_layers(ns):
    g = torch.Generator().manual_seed(2)
    act = torch.nn.SiLU()
    arrs = {}
    for name, kw, cin in [("plain", dict(in_ch=8, out_ch=8), 8), ("chg", dict(in_ch=8, out_ch=16), 8),
                          ("down", dict(in_ch=8, out_ch=8, down=True), 8), ("up", dict(in_ch=8, out_ch=8, up=True), 8),
                          ("cat", dict(in_ch=24, out_ch=16), 24)]:
        blk = ns.layerspp.ResnetBlockBigGANpp(act=act, dropout=0.0, fir=True, fir_kernel=[1, 3, 3, 1], init_scale=1.0,
                                              skip_rescale=True, temb_dim=32, **kw).eval()
        randomize_(blk, g)
        x = torch.randn(2, cin, 8, 12, generator=g)
        temb = torch.randn(2, 32, generator=g)
        with torch.no_grad():
            y = blk(x, temb)
        arrs.update({f"{name}/x": x, f"{name}/temb": temb, f"{name}/y": y})
        arrs.update(sd_arrays(blk.state_dict(), f"{name}/sd/blk."))
    attn = ns.layerspp.AttnBlockpp(channels=16, skip_rescale=True, init_scale=1.0).eval()
    randomize_(attn, g)
    x = torch.randn(2, 16, 6, 5, generator=g)
    with torch.no_grad():
        y = attn(x)
    arrs.update({"attn/x": x, "attn/y": y})
    arrs.update(sd_arrays(attn.state_dict(), "attn/sd/blk."))
    save("layers", **arrs)


def _tiny_model(ns, name, gen, **kw):
    cls = ns.BackboneRegistry.get_by_name(name)
    m = cls(init_scale=1.0, **kw).eval()
    randomize_(m, gen, scale=0.15)
    with torch.no_grad():
        m.all_modules[0].W.copy_(torch.randn(m.all_modules[0].W.shape, generator=gen) * 16.0)
    return m


def gen_net_tiny(ns):
    g = torch.Generator().manual_seed(3)
    m = _tiny_model(ns, "ncsnpp", g, nf=8, input_channels=4)
    x = cplx(g, 2, 2, 16, 24, scale=0.5)
    t = torch.tensor([0.8, 0.31])
    with torch.no_grad():
        y = m(x, t)
    save("ncsnpp_tiny", x=x, t=t, y=y, **sd_arrays(m.state_dict()))
    # discriminative (denoiser) variant: 2 input channels, no time conditioning (ncsnpp.py:80-86)
    md = _tiny_model(ns, "ncsnpp", g, nf=8, discriminative=True)
    xd = cplx(g, 2, 1, 16, 24, scale=0.5)
    with torch.no_grad():
        yd = md(xd)
    save("ncsnpp_tiny_disc", x=xd, y=yd, **sd_arrays(md.state_dict()))
    # 6-channel score net as used by StochasticRegenerationModel (model.py:421)
    m6 = _tiny_model(ns, "ncsnpp", g, nf=8, input_channels=6)
    x6 = cplx(g, 1, 3, 16, 24, scale=0.5)
    t6 = torch.tensor([0.5])
    with torch.no_grad():
        y6 = m6(x6, t6)
    save("ncsnpp_tiny_6ch", x=x6, t=t6, y=y6, **sd_arrays(m6.state_dict()))


def gen_net_large_tiny(ns):
    g = torch.Generator().manual_seed(4)
    m = _tiny_model(ns, "ncsnpplarge", g, input_channels=4)
    # NCSNppLarge hard-codes nf=128; build the same topology at nf=8 through the base class
    m = ns.BackboneRegistry.get_by_name("ncsnpp")(nf=8, ch_mult=(1, 1, 2, 2, 2, 2, 2), num_res_blocks=2,
                                                   attn_resolutions=(16,), init_scale=1.0, input_channels=4).eval()
    randomize_(m, g, scale=0.15)
    with torch.no_grad():
        m.all_modules[0].W.copy_(torch.randn(m.all_modules[0].W.shape, generator=g) * 16.0)
    x = cplx(g, 1, 2, 256, 64, scale=0.5)
    t = torch.tensor([0.63])
    with torch.no_grad():
        y = m(x, t)
    save("ncsnpplarge_tiny", x=x, t=t, y=y, **sd_arrays(m.state_dict()))


def gen_net_full(ns):
    """Full-width ncsnpp / ncsnpplarge with weights regenerated from a seed at test time (not stored)."""
    g = torch.Generator().manual_seed(5)
    cfg = ncsnpp_spec.NCSNppConfig()
    sd = ncsnpp_spec.random_state_dict(cfg, seed=0, init_scale=1.0, bias_std=0.05)
    m = ns.BackboneRegistry.get_by_name("ncsnpp")(input_channels=4, init_scale=1.0).eval()
    m.load_state_dict(sd, strict=True)
    x = cplx(g, 2, 2, 64, 64, scale=0.5)
    t = torch.tensor([1.0, 0.2])
    with torch.no_grad():
        y = m(x, t)
    save("ncsnpp_full_64x64", x=x, t=t, y=y, seed=0, init_scale=1.0, bias_std=0.05)

    cfgl = ncsnpp_spec.NCSNppConfig(ch_mult=(1, 1, 2, 2, 2, 2, 2), num_res_blocks=2, attn_resolutions=(16,))
    sdl = ncsnpp_spec.random_state_dict(cfgl, seed=1, init_scale=1.0, bias_std=0.05)
    ml = ns.BackboneRegistry.get_by_name("ncsnpplarge")(input_channels=4, init_scale=1.0).eval()
    ml.load_state_dict(sdl, strict=True)
    xl = cplx(g, 1, 2, 256, 64, scale=0.5)
    tl = torch.tensor([0.47])
    with torch.no_grad():
        yl = ml(xl, tl)
    save("ncsnpplarge_full_256x64", x=xl, t=tl, y=yl, seed=1, init_scale=1.0, bias_std=0.05)


def _score_fn(m):
    def score(x, t, y, **kw):           # ScoreModel.forward, model.py:127-136
        return -m(torch.cat([x, y], dim=1), t)
    return score


def gen_sampler(ns):
    g = torch.Generator().manual_seed(6)
    m = _tiny_model(ns, "ncsnpp", g, nf=8, input_channels=4)
    y = cplx(g, 2, 1, 16, 64, scale=0.3)
    arrs = dict(y=y, **sd_arrays(m.state_dict()))
    N = 3
    for tag, pred, corr, csteps, snr in [("rd_ald", "reverse_diffusion", "ald", 1, 0.5),
                                         ("rd_ald2", "reverse_diffusion", "ald", 2, 0.33),
                                         ("em_langevin", "euler_maruyama", "langevin", 1, 0.5),
                                         ("rd_none", "reverse_diffusion", "none", 1, 0.5),
                                         ("none_ald", "none", "ald", 1, 0.5)]:
        ndraw = 1 + N * ((0 if corr == "none" else csteps) + (0 if pred == "none" else 1))
        zs = [cplx(g, *y.shape) * float(np.sqrt(0.5)) for _ in range(ndraw)]
        sde = ns.OUVESDE(theta=1.5, sigma_min=0.05, sigma_max=0.5, N=N)
        sampler = ns.sampling.get_pc_sampler(pred, corr, sde=sde, score_fn=_score_fn(m), y=y, eps=0.03, snr=snr,
                                             corrector_steps=csteps)
        q = NoiseQueue(zs)
        with patched_randn_like(q):
            x, nfe = sampler()
        assert q.i == ndraw, (tag, q.i, ndraw)
        arrs.update({f"{tag}/noise": torch.stack(zs), f"{tag}/x": x, f"{tag}/nfe": nfe, f"{tag}/N": N,
                     f"{tag}/snr": snr, f"{tag}/csteps": csteps})
    # SDE scalars
    sde = ns.OUVESDE(theta=1.5, sigma_min=0.05, sigma_max=0.5, N=30)
    tt = torch.linspace(1.0, 0.03, 30)
    arrs["sde/t"] = tt
    arrs["sde/std"] = sde._std(tt)
    arrs["sde/diffusion"] = sde.sde(torch.zeros(30, 1, 1, 1, dtype=torch.complex64), tt, torch.zeros(30, 1, 1, 1, dtype=torch.complex64))[1]
    # ODE sampler (scipy RK45 on host, sampling/__init__.py:71-141)
    zs = [cplx(g, *y.shape) * float(np.sqrt(0.5)) for _ in range(2)]   # prior + the (discarded) denoise-step draw
    sde = ns.OUVESDE(theta=1.5, sigma_min=0.05, sigma_max=0.5, N=5)
    ode = ns.sampling.get_ode_sampler(sde, _score_fn(m), y=y, eps=0.03, device="cpu")
    q = NoiseQueue(zs)
    with patched_randn_like(q):
        x, nfe = ode()
    arrs.update({"ode/noise": torch.stack(zs), "ode/x": x, "ode/nfe": nfe, "ode/N": 5})
    save("sampler", **arrs)


def gen_stft(ns):
    g = torch.Generator().manual_seed(7)
    arrs = {}
    for tag, L in [("a", 5000), ("b", 32640), ("c", 1000)]:
        sig = torch.randn(1, L, generator=g) * 0.1
        dm = ns.SpecsDataModule(spec_factor=0.15, spec_abs_exponent=0.5)
        S = dm.stft(sig)
        Sf = dm.spec_fwd(S)
        Sp = ns.pad_spec(Sf.unsqueeze(0))
        back = dm.istft(dm.spec_back(Sf), L)
        arrs.update({f"{tag}/sig": sig, f"{tag}/stft": S, f"{tag}/fwd": Sf, f"{tag}/padded_T": Sp.shape[-1], f"{tag}/back": back})
    dm2 = ns.SpecsDataModule(spec_factor=0.33, spec_abs_exponent=1.0, window="sqrthann")
    sig = torch.randn(1, 3000, generator=g) * 0.1
    S = dm2.stft(sig)
    arrs.update({"d/sig": sig, "d/stft": S, "d/fwd": dm2.spec_fwd(S), "d/back": dm2.istft(dm2.spec_back(dm2.spec_fwd(S)), 3000)})
    save("stft", **arrs)


def gen_enhance(ns):
    """End-to-end glue of ScoreModel.enhance (model.py:282-303) around reference modules, minus .cuda()."""
    g = torch.Generator().manual_seed(8)
    m = _tiny_model(ns, "ncsnpp", g, nf=8, input_channels=4)
    dm = ns.SpecsDataModule(spec_factor=0.15, spec_abs_exponent=0.5)
    yw = torch.randn(1, 9000, generator=g) * 0.2
    T_orig = yw.size(1)
    norm = yw.abs().max().item()
    Y = ns.pad_spec(torch.unsqueeze(dm.spec_fwd(dm.stft(yw / norm)), 0))
    N = 2
    zs = [cplx(g, *Y.shape) * float(np.sqrt(0.5)) for _ in range(1 + 2 * N)]
    sde = ns.OUVESDE(theta=1.5, sigma_min=0.05, sigma_max=0.5, N=N)
    sampler = ns.sampling.get_pc_sampler("reverse_diffusion", "ald", sde=sde, score_fn=_score_fn(m), y=Y, eps=0.03,
                                         snr=0.5, corrector_steps=1)
    with patched_randn_like(NoiseQueue(zs)):
        sample, nfe = sampler()
    x_hat = dm.istft(dm.spec_back(sample.squeeze()), T_orig) * norm
    save("enhance", y=yw, noise=torch.stack(zs), Y=Y, sample=sample, x_hat=x_hat.squeeze(), nfe=nfe, N=N,
         **sd_arrays(m.state_dict()))


def gen_storm(ns):
    """StochasticRegenerationModel.enhance glue (model.py:720-780) around reference modules, condition="both"."""
    g = torch.Generator().manual_seed(9)
    den = _tiny_model(ns, "ncsnpp", g, nf=8, discriminative=True)
    score = _tiny_model(ns, "ncsnpp", g, nf=8, input_channels=6)
    dm = ns.SpecsDataModule(spec_factor=0.15, spec_abs_exponent=0.5)
    yw = torch.randn(1, 6000, generator=g) * 0.2
    T_orig = yw.size(1)
    norm = yw.abs().max().item()
    Y = ns.pad_spec(torch.unsqueeze(dm.spec_fwd(dm.stft(yw / norm)), 0))
    with torch.no_grad():
        Y_den = den(Y)
    cond = [Y, Y_den]

    def forward_score(x, t, score_conditioning=None, sde_input=None, **kw):       # model.py:654-661
        return -score(torch.cat([x] + score_conditioning, dim=1), t)

    N = 3
    zs = [cplx(g, *Y.shape) * float(np.sqrt(0.5)) for _ in range(1 + N)]
    sde = ns.OUVESDE(theta=1.5, sigma_min=0.05, sigma_max=0.5, N=N)
    sampler = ns.sampling.get_pc_sampler("reverse_diffusion", "none", sde=sde, score_fn=forward_score, y=Y_den, eps=0.03, snr=0.5,
                                         corrector_steps=1, conditioning=cond)
    with patched_randn_like(NoiseQueue(zs)):
        sample, nfe = sampler()
    x_hat = dm.istft(dm.spec_back(sample.squeeze()), T_orig) * norm
    save("storm", y=yw, noise=torch.stack(zs), Y=Y, Y_den=Y_den, sample=sample, x_hat=x_hat.squeeze(), nfe=nfe, N=N,
         **sd_arrays(den.state_dict(), "den/"), **sd_arrays(score.state_dict(), "score/"))


def gen_headline(ns):
    """The headline benchmark configuration end to end through the UNMODIFIED reference modules (one utterance of
    BASELINE configs[1]): 32 640 samples -> 256x256, full-width ncsnpp with seeded weights, N=30 reverse_diffusion +
    ald x1 (60 network evaluations), noise from oracle.seeded_complex_noise.  Only the outputs are stored; weights,
    waveform and noise are regenerated from their seeds at test time.  Takes a few minutes on CPU."""
    from oracle import storm_oracle as O
    cfg = ncsnpp_spec.NCSNppConfig()
    sd = ncsnpp_spec.random_state_dict(cfg, seed=0, init_scale=1.0, bias_std=0.05)
    m = ns.BackboneRegistry.get_by_name("ncsnpp")(input_channels=4, init_scale=1.0).eval()
    m.load_state_dict(sd, strict=True)
    dm = ns.SpecsDataModule(spec_factor=0.15, spec_abs_exponent=0.5)
    yw = O.synth_utterance(0, 32640).reshape(1, -1)
    T_orig = yw.size(1)
    norm = yw.abs().max().item()
    Y = ns.pad_spec(torch.unsqueeze(dm.spec_fwd(dm.stft(yw / norm)), 0))
    assert tuple(Y.shape) == (1, 1, 256, 256)
    N = 30
    zs = O.seeded_complex_noise(2024, 1 + 2 * N, tuple(Y.shape))
    sde = ns.OUVESDE(theta=1.5, sigma_min=0.05, sigma_max=0.5, N=N)
    sampler = ns.sampling.get_pc_sampler("reverse_diffusion", "ald", sde=sde, score_fn=_score_fn(m), y=Y, eps=0.03,
                                         snr=0.5, corrector_steps=1)
    with patched_randn_like(NoiseQueue(zs)), torch.no_grad():
        sample, nfe = sampler()
    x_hat = dm.istft(dm.spec_back(sample.squeeze()), T_orig) * norm
    save("headline_256x256_N30", sample=sample, x_hat=x_hat.squeeze(), nfe=nfe, N=N, noise_seed=2024, weight_seed=0,
         init_scale=1.0, bias_std=0.05, L=T_orig)


def gen_metrics(ns):
    """si_sdr / si_sdr_torch (util/other.py:82-94) on clean/estimate pairs at several SNRs and lengths."""
    g = torch.Generator().manual_seed(10)
    arrs = {}
    for tag, L, noise in [("a", 4000, 0.5), ("b", 32640, 0.05), ("c", 16000, 0.001), ("d", 777, 2.0)]:
        s = torch.randn(L, generator=g) * 0.1
        s_hat = 0.7 * s + noise * 0.1 * torch.randn(L, generator=g)
        arrs.update({f"{tag}/s": s, f"{tag}/s_hat": s_hat,
                     f"{tag}/si_sdr": np.float64(ns.si_sdr(s.double().numpy(), s_hat.double().numpy())),
                     f"{tag}/si_sdr_torch": ns.si_sdr_torch(s, s_hat)})
    s = torch.randn(3000, generator=g) * 0.1
    arrs.update({"e/s": s, "e/s_hat": s[:2500] * 1.3 + 0.01 * torch.randn(2500, generator=g)})
    arrs["e/si_sdr_torch"] = ns.si_sdr_torch(arrs["e/s"], arrs["e/s_hat"])      # different lengths: truncation
    save("metrics", **arrs)


def gen_param_order(ns):
    """``named_parameters()`` of the reference networks: name, shape (at nf=8) and requires_grad, in the order
    ``torch_ema`` sees them (sgmse/model.py:55,102-107: ExponentialMovingAverage(self.parameters())).  Checkpoint loading
    maps ``ema.shadow_params`` positionally onto this list, so the order and the one frozen parameter
    (GaussianFourierProjection.W, layerspp.py:37) are pinned here."""
    import json
    N = ns.BackboneRegistry.get_by_name("ncsnpp")
    out = {}
    for tag, kw in [("ncsnpp_4ch", dict(input_channels=4)), ("ncsnpp_6ch", dict(input_channels=6)),
                    ("ncsnpp_disc", dict(input_channels=2, discriminative=True))]:
        m = N(nf=8, **kw)
        assert [n for n, _ in m.named_parameters()] == list(m.state_dict().keys())
        out[tag] = [[n, list(p.shape), bool(p.requires_grad)] for n, p in m.named_parameters()]
    path = os.path.join(OUT, "param_order.json")
    with open(path, "w") as f:
        json.dump(out, f)
    print(f"wrote {path}")


GENS = dict(param_order=gen_param_order, headline=gen_headline, metrics=gen_metrics, storm=gen_storm, fir=gen_fir, layers=gen_layers, net_tiny=gen_net_tiny, net_large_tiny=gen_net_large_tiny,
            net_full=gen_net_full, sampler=gen_sampler, stft=gen_stft, enhance=gen_enhance)

if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--only", default=None)
    a = ap.parse_args()
    torch.set_num_threads(os.cpu_count())
    os.makedirs(OUT, exist_ok=True)
    ns = ref_shim.load()
    for k, fn in GENS.items():
        if a.only in (None, k):
            fn(ns)
