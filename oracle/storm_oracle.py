"""CPU oracle for the reverse-diffusion enhancement path of sp-uhh/storm.

TEST INFRASTRUCTURE ONLY.  Nothing under ``storm_b200/`` may import this file; only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs do, and there
only as the checker or the timed CPU baseline.

It is a from-scratch functional restatement (plain torch fp32 ops on CPU, no nn.Module tree) of what the
reference computes on the hot path; every function cites the reference ``file:line`` it follows
(paths relative to ``/root/reference``).  The reference ships no tests or golden vectors (SURVEY.md
section 4), so the restatement is pinned against the *reference itself executed in the build container*:
``oracle/gen_golden.py`` imports the unmodified reference (``oracle/ref_shim.py``), runs it on seeded
inputs and commits inputs/outputs under ``tests/golden/``; ``tests/test_oracle_golden.py`` checks this
file against those fixtures.  Third-party arithmetic (conv2d, group_norm, softmax, stft, istft) lives in
PyTorch (reference pins torch==1.11.0, ``requirements.txt:14``; here torch 2.11 CPU kernels).

All weights are read from a ``state_dict`` that is key-compatible with the reference
(``all_modules.{i}.*``, ``output_layer.*``).
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Callable, Dict, List, Optional, Tuple

import numpy as np
import torch
import torch.nn.functional as F

Tensor = torch.Tensor
SQRT2 = float(np.sqrt(2.0))


# --------------------------------------------------------------------------------------------------
# configuration
# --------------------------------------------------------------------------------------------------
@dataclass
class NetCfg:
    """Subset of NCSNpp ctor options (sgmse/backbones/ncsnpp.py:40-65) that the path uses."""

    nf: int = 128
    ch_mult: Tuple[int, ...] = (1, 2, 2, 2)
    num_res_blocks: int = 1
    attn_resolutions: Tuple[int, ...] = (0,)
    image_size: int = 256
    input_channels: int = 4
    discriminative: bool = False
    fourier_scale: float = 16.0

    @property
    def conditional(self) -> bool:  # ncsnpp.py:81-86
        return not self.discriminative

    @property
    def scale_by_sigma(self) -> bool:
        return not self.discriminative

    @property
    def in_ch(self) -> int:
        return 2 if self.discriminative else self.input_channels


NAMED_CFGS = {
    "ncsnpp": dict(),                                                        # ncsnpp.py:36-65
    "ncsnpplarge": dict(ch_mult=(1, 1, 2, 2, 2, 2, 2), num_res_blocks=2, attn_resolutions=(16,)),  # :460-470
    "ncsnpp12M": dict(nf=96, ch_mult=(1, 2, 2, 1)),                          # :479-489
    "ncsnpp6M": dict(nf=96, ch_mult=(1, 1, 1, 1)),                           # :498-508
}


def named_cfg(name: str, **over) -> NetCfg:
    kw = dict(NAMED_CFGS[name])
    kw.update(over)
    return NetCfg(**kw)


# --------------------------------------------------------------------------------------------------
# FIR resampling: closed forms of upfirdn2d with taps [1,3,3,1]
# (up_or_down_sampling.py:195-257, op/upfirdn2d.py:161-202)
# --------------------------------------------------------------------------------------------------
def _fir_down_axis(x: Tensor, dim: int) -> Tensor:
    """o[i] = (x[2i-1] + 3 x[2i] + 3 x[2i+1] + x[2i+2]) / 8, zeros outside."""
    n = x.shape[dim]
    pad = [0, 0] * x.dim()
    # F.pad pads last dim first
    idx = 2 * (x.dim() - 1 - dim)
    pad[idx] = 1
    pad[idx + 1] = 1
    xp = F.pad(x, pad)
    a = xp.narrow(dim, 0, n).unfold(dim, 1, 2).squeeze(-1)      # x[2i-1]
    b = xp.narrow(dim, 1, n).unfold(dim, 1, 2).squeeze(-1)      # x[2i]
    c = xp.narrow(dim, 2, n).unfold(dim, 1, 2).squeeze(-1)      # x[2i+1]
    d = xp.narrow(dim, 3, n - 1).unfold(dim, 1, 2).squeeze(-1)  # x[2i+2]
    return (a + 3.0 * b + 3.0 * c + d) * 0.125


def fir_downsample(x: Tensor) -> Tensor:
    """downsample_2d(x, [1,3,3,1], factor=2) on [B,C,H,W] (up_or_down_sampling.py:227-257)."""
    return _fir_down_axis(_fir_down_axis(x, 2), 3)


def _fir_up_axis(x: Tensor, dim: int) -> Tensor:
    """o[2i] = x[i-1]/4 + 3x[i]/4 ; o[2i+1] = 3x[i]/4 + x[i+1]/4, zeros outside."""
    n = x.shape[dim]
    pad = [0, 0] * x.dim()
    idx = 2 * (x.dim() - 1 - dim)
    pad[idx] = 1
    pad[idx + 1] = 1
    xp = F.pad(x, pad)
    prev = xp.narrow(dim, 0, n)
    cur = xp.narrow(dim, 1, n)
    nxt = xp.narrow(dim, 2, n)
    even = 0.25 * prev + 0.75 * cur
    odd = 0.75 * cur + 0.25 * nxt
    out = torch.stack([even, odd], dim=dim + 1)
    shp = list(x.shape)
    shp[dim] = 2 * n
    return out.reshape(shp)


def fir_upsample(x: Tensor) -> Tensor:
    """upsample_2d(x, [1,3,3,1], factor=2) on [B,C,H,W] (up_or_down_sampling.py:195-224)."""
    return _fir_up_axis(_fir_up_axis(x, 2), 3)


# --------------------------------------------------------------------------------------------------
# NCSN++ building blocks
# --------------------------------------------------------------------------------------------------
def _gn(x: Tensor, sd, key: str) -> Tensor:
    c = x.shape[1]
    return F.group_norm(x, min(c // 4, 32), sd[key + ".weight"], sd[key + ".bias"], eps=1e-6)


def _conv(x: Tensor, sd, key: str, pad: int) -> Tensor:
    return F.conv2d(x, sd[key + ".weight"], sd[key + ".bias"], padding=pad)


def _nin(x: Tensor, sd, key: str) -> Tensor:
    """NIN: per-pixel x @ W + b with W [in, out] (layers.py:548-557)."""
    w = sd[key + ".W"]
    return torch.einsum("bchw,cd->bdhw", x, w) + sd[key + ".b"][None, :, None, None]


def time_embedding(t: Tensor, sd, cfg: NetCfg) -> Optional[Tensor]:
    """Gaussian-Fourier features of log t then Linear-SiLU-Linear (ncsnpp.py:298-317, layerspp.py:39-41)."""
    if not cfg.conditional:
        return None
    w = sd["all_modules.0.W"]
    x_proj = torch.log(t)[:, None] * w[None, :] * 2 * np.pi
    e = torch.cat([torch.sin(x_proj), torch.cos(x_proj)], dim=-1)
    h = F.linear(e, sd["all_modules.1.weight"], sd["all_modules.1.bias"])
    return F.linear(F.silu(h), sd["all_modules.2.weight"], sd["all_modules.2.bias"])


def resblock(x: Tensor, temb: Optional[Tensor], sd, p: str, up: bool = False, down: bool = False) -> Tensor:
    """ResnetBlockBigGANpp.forward (layerspp.py:242-274)."""
    h = F.silu(_gn(x, sd, p + ".GroupNorm_0"))
    if up:
        h, x = fir_upsample(h), fir_upsample(x)
    elif down:
        h, x = fir_downsample(h), fir_downsample(x)
    h = _conv(h, sd, p + ".Conv_0", 1)
    if temb is not None:
        h = h + F.linear(F.silu(temb), sd[p + ".Dense_0.weight"], sd[p + ".Dense_0.bias"])[:, :, None, None]
    h = F.silu(_gn(h, sd, p + ".GroupNorm_1"))
    h = _conv(h, sd, p + ".Conv_1", 1)
    if (p + ".Conv_2.weight") in sd:
        x = _conv(x, sd, p + ".Conv_2", 0)
    return (x + h) / SQRT2


def attnblock(x: Tensor, sd, p: str) -> Tensor:
    """AttnBlockpp.forward: single-head attention over all H*W positions (layerspp.py:75-91)."""
    b, c, hh, ww = x.shape
    h = _gn(x, sd, p + ".GroupNorm_0")
    q = _nin(h, sd, p + ".NIN_0").reshape(b, c, hh * ww)
    k = _nin(h, sd, p + ".NIN_1").reshape(b, c, hh * ww)
    v = _nin(h, sd, p + ".NIN_2").reshape(b, c, hh * ww)
    w = torch.einsum("bcq,bck->bqk", q, k) * (int(c) ** (-0.5))
    w = torch.softmax(w, dim=-1)
    h = torch.einsum("bqk,bck->bcq", w, v).reshape(b, c, hh, ww)
    h = _nin(h, sd, p + ".NIN_3")
    return (x + h) / SQRT2


def ncsnpp_forward(sd: Dict[str, Tensor], cfg: NetCfg, x: Tensor, t: Optional[Tensor],
                   taps: Optional[dict] = None) -> Tensor:
    """NCSNpp.forward (ncsnpp.py:281-450) for the default topology
    (biggan blocks, fir, output_skip / input_skip / sum, skip_rescale, fourier embedding).

    x: complex64 [B, in_ch/2, F, T];  t: float32 [B] (None when discriminative).
    ``taps``: optional dict that receives named intermediate activations (for layer-level parity tests).
    """
    nres = len(cfg.ch_mult)
    # complex -> real channels [re0, im0, re1, im1, ...]   (ncsnpp.py:290-296)
    h0 = torch.cat([torch.cat([x[:, [i]].real, x[:, [i]].imag], dim=1) for i in range(cfg.in_ch // 2)], dim=1)
    temb = time_embedding(t, sd, cfg)
    m = 3 if cfg.conditional else 1
    h0 = 2 * h0 - 1.0                                        # ncsnpp.py:321-323
    pyr_in = h0
    hs = [_conv(h0, sd, f"all_modules.{m}", 1)]              # :330
    m += 1

    def tap(name, v):
        if taps is not None:
            taps[name] = v

    tap("conv_in", hs[0])
    for lvl in range(nres):                                  # :332-366
        for _ in range(cfg.num_res_blocks):
            h = resblock(hs[-1], temb, sd, f"all_modules.{m}")
            tap(f"mod{m}", h)
            m += 1
            if h.shape[-2] in cfg.attn_resolutions:
                h = attnblock(h, sd, f"all_modules.{m}")
                tap(f"mod{m}", h)
                m += 1
            hs.append(h)
        if lvl != nres - 1:
            h = resblock(hs[-1], temb, sd, f"all_modules.{m}", down=True)
            tap(f"mod{m}", h)
            m += 1
            pyr_in = fir_downsample(pyr_in)                  # :353
            h = _conv(pyr_in, sd, f"all_modules.{m}.Conv_0", 0) + h   # Combine 'sum' (layerspp.py:52-57)
            tap(f"mod{m}", h)
            m += 1
            hs.append(h)

    h = hs[-1]
    h = resblock(h, temb, sd, f"all_modules.{m}"); tap(f"mod{m}", h); m += 1      # :369
    h = attnblock(h, sd, f"all_modules.{m}"); tap(f"mod{m}", h); m += 1           # :371
    h = resblock(h, temb, sd, f"all_modules.{m}"); tap(f"mod{m}", h); m += 1      # :373

    pyr = None
    for lvl in reversed(range(nres)):                        # :379-428
        for _ in range(cfg.num_res_blocks + 1):
            h = resblock(torch.cat([h, hs.pop()], dim=1), temb, sd, f"all_modules.{m}")
            tap(f"mod{m}", h)
            m += 1
        if h.shape[-2] in cfg.attn_resolutions:
            h = attnblock(h, sd, f"all_modules.{m}")
            tap(f"mod{m}", h)
            m += 1
        ph = F.silu(_gn(h, sd, f"all_modules.{m}"))          # :392 / :406
        m += 1
        ph = _conv(ph, sd, f"all_modules.{m}", 1)            # :394 / :408
        m += 1
        pyr = ph if pyr is None else fir_upsample(pyr) + ph  # :405,410
        tap(f"pyr{lvl}", pyr)
        if lvl != 0:
            h = resblock(h, temb, sd, f"all_modules.{m}", up=True)   # :427
            tap(f"mod{m}", h)
            m += 1
    assert not hs
    h = pyr                                                  # :432-433
    if cfg.scale_by_sigma:
        h = h / t[:, None, None, None]                       # :441-443
    h = F.conv2d(h, sd["output_layer.weight"], sd["output_layer.bias"])   # :446
    return torch.complex(h[:, 0:1].contiguous(), h[:, 1:2].contiguous())  # :447-449 (spatial_channels=1)


def count_modules(cfg: NetCfg) -> int:
    """Length of ``all_modules`` for this config (ncsnpp.py:110-273)."""
    nres = len(cfg.ch_mult)
    res = [cfg.image_size // (2 ** i) for i in range(nres)]
    n = 3 if cfg.conditional else 1
    n += 1
    for lvl in range(nres):
        for _ in range(cfg.num_res_blocks):
            n += 1 + (1 if res[lvl] in cfg.attn_resolutions else 0)
        if lvl != nres - 1:
            n += 2
    n += 3
    for lvl in reversed(range(nres)):
        n += cfg.num_res_blocks + 1
        n += 1 if res[lvl] in cfg.attn_resolutions else 0
        n += 2
        n += 1 if lvl != 0 else 0
    return n


# --------------------------------------------------------------------------------------------------
# OUVE SDE + samplers
# --------------------------------------------------------------------------------------------------
@dataclass
class OUVE:
    """OUVESDE constants (sdes.py:166-252)."""

    theta: float = 1.5
    sigma_min: float = 0.05
    sigma_max: float = 0.5

    @property
    def logsig(self) -> float:
        return float(np.log(self.sigma_max / self.sigma_min))

    def diffusion(self, t: Tensor) -> Tensor:                      # sdes.py:200-208
        sigma = self.sigma_min * (self.sigma_max / self.sigma_min) ** t
        return sigma * np.sqrt(2 * self.logsig)

    def std(self, t: Tensor) -> Tensor:                            # sdes.py:215-228
        th, ls, sm = self.theta, self.logsig, self.sigma_min
        return torch.sqrt((sm ** 2 * torch.exp(-2 * th * t) * (torch.exp(2 * (th + ls) * t) - 1) * ls) / (th + ls))


NoiseFn = Callable[[Tensor], Tensor]


def _default_noise(x: Tensor) -> Tensor:
    return torch.randn_like(x)


def pc_sample(score_fn: Callable[[Tensor, Tensor, Tensor], Tensor], y: Tensor, sde: OUVE, N: int,
              predictor: str = "reverse_diffusion", corrector: str = "ald", corrector_steps: int = 1,
              snr: float = 0.5, eps: float = 3e-2, noise: NoiseFn = _default_noise,
              denoise: bool = True, probability_flow_discretize: bool = False) -> Tuple[Tensor, int]:
    """Predictor-corrector sampler (sampling/__init__.py:54-66) with
    ReverseDiffusion / EulerMaruyama / none predictors (predictors.py:41-80) and
    ald / langevin / none correctors (correctors.py:37-106).

    ``noise(x)`` replaces ``torch.randn_like(x)``; draws happen in the reference's order
    (prior, then per step: corrector draws, predictor draw).
    ``probability_flow_discretize`` selects RSDE.discretize's PF branch (sdes.py:153-156), which the
    reference's own PC sampler never enables (predictors.py:15-20 drops the flag) - used for config 5.
    """
    B = y.shape[0]
    ones = torch.ones(B, dtype=torch.float32)
    xt = y + noise(y) * sde.std(ones)[:, None, None, None]          # prior_sampling, sdes.py:233-237
    ts = torch.linspace(1.0, eps, N)
    xm = xt
    n_corr = 0 if corrector == "none" else corrector_steps
    for i in range(N):
        t = ones * ts[i]
        # ---- corrector
        if corrector == "ald":                                       # correctors.py:76-93
            std = sde.std(t)
            for _ in range(corrector_steps):
                g = score_fn(xt, t, y)
                z = noise(xt)
                step = ((snr * std) ** 2 * 2)[:, None, None, None]
                xm = xt + step * g
                xt = xm + z * torch.sqrt(step * 2)
        elif corrector == "langevin":                                # correctors.py:45-61
            for _ in range(corrector_steps):
                g = score_fn(xt, t, y)
                z = noise(xt)
                gn = torch.norm(g.reshape(B, -1), dim=-1).mean()
                zn = torch.norm(z.reshape(B, -1), dim=-1).mean()
                step = (snr * zn / gn) ** 2 * 2
                xm = xt + step * g
                xt = xm + z * torch.sqrt(step * 2)
        elif corrector != "none":
            raise ValueError(f"Corrector with name '{corrector}' unknown.")
        # ---- predictor
        if predictor == "reverse_diffusion":                         # predictors.py:62-69, sdes.py:73-90,147-157
            dt = 1.0 / N
            f = sde.theta * (y - xt) * dt
            G = (sde.diffusion(t) * torch.sqrt(torch.tensor(dt)))[:, None, None, None]
            s = score_fn(xt, t, y)
            if probability_flow_discretize:
                rev_f, rev_G = f - G ** 2 * s * 0.5, torch.zeros_like(G)
            else:
                rev_f, rev_G = f - G ** 2 * s, G
            z = noise(xt)
            xm = xt - rev_f
            xt = xm + rev_G * z
        elif predictor == "euler_maruyama":                          # predictors.py:46-54, sdes.py:117-145
            dt = -1.0 / N
            z = noise(xt)
            g = sde.diffusion(t)[:, None, None, None]
            drift = sde.theta * (y - xt) - g ** 2 * score_fn(xt, t, y)
            xm = xt + drift * dt
            xt = xm + g * np.sqrt(-dt) * z
        elif predictor == "none":
            xm = xt
        else:
            raise ValueError(f"Predictor with name '{predictor}' unknown.")
    return (xm if (denoise and N) else xt), N * (n_corr + 1)


def ode_drift(score_fn, sde: OUVE, x: Tensor, t: Tensor, y: Tensor) -> Tensor:
    """Probability-flow drift  theta (y - x) - 1/2 g(t)^2 score  (sdes.py:124-141, probability_flow=True)."""
    g = sde.diffusion(t)[:, None, None, None]
    return sde.theta * (y - x) - 0.5 * g ** 2 * score_fn(x, t, y)


def ode_sample(score_fn, y: Tensor, sde: OUVE, N: int, eps: float = 3e-2, rtol: float = 1e-5, atol: float = 1e-5,
               method: str = "RK45", noise: NoiseFn = _default_noise, denoise: bool = True) -> Tuple[Tensor, int]:
    """get_ode_sampler (sampling/__init__.py:71-141): scipy RK45 on the flattened complex state, then one
    noise-free reverse-diffusion step at t=eps with dt=1/N (:97-100,134-135)."""
    from scipy import integrate

    B = y.shape[0]
    ones = torch.ones(B, dtype=torch.float32)
    x = y + noise(y) * sde.std(ones)[:, None, None, None]

    def rhs(tt, xf):
        xx = torch.from_numpy(xf.reshape(tuple(y.shape))).type(torch.complex64)
        return ode_drift(score_fn, sde, xx, ones * tt, y).detach().cpu().numpy().reshape((-1,))

    sol = integrate.solve_ivp(rhs, (1.0, eps), x.detach().cpu().numpy().reshape((-1,)), rtol=rtol, atol=atol, method=method)
    x = torch.tensor(sol.y[:, -1]).reshape(y.shape).type(torch.complex64)
    nfe = sol.nfev
    if denoise:
        t = ones * eps
        dt = 1.0 / N
        f = sde.theta * (y - x) * dt
        G = (sde.diffusion(t) * torch.sqrt(torch.tensor(dt)))[:, None, None, None]
        x = x - (f - G ** 2 * score_fn(x, t, y))
    return x, nfe


def make_score_fn(sd, cfg: NetCfg) -> Callable[[Tensor, Tensor, Tensor], Tensor]:
    """ScoreModel.forward: score = -dnn(cat[x, y], t)   (model.py:127-136)."""

    def score(x, t, y):
        return -ncsnpp_forward(sd, cfg, torch.cat([x, y], dim=1), t)

    return score


# --------------------------------------------------------------------------------------------------
# STFT front/back end  (data_module.py:182-223, util/other.py:102-109)
# --------------------------------------------------------------------------------------------------
@dataclass
class SpecCfg:
    n_fft: int = 510
    hop_length: int = 128
    spec_factor: float = 0.15
    spec_abs_exponent: float = 0.5
    window: str = "hann"

    def win(self) -> Tensor:                                        # data_module.py:19-25
        w = torch.hann_window(self.n_fft, periodic=True)
        return torch.sqrt(w) if self.window == "sqrthann" else w


def stft(sig: Tensor, sc: SpecCfg) -> Tensor:                       # data_module.py:217-219
    return torch.stft(sig, n_fft=sc.n_fft, hop_length=sc.hop_length, window=sc.win(), center=True, return_complex=True)


def istft(spec: Tensor, sc: SpecCfg, length: Optional[int] = None) -> Tensor:   # data_module.py:221-223
    return torch.istft(spec, n_fft=sc.n_fft, hop_length=sc.hop_length, window=sc.win(), center=True, length=length)


def spec_fwd(spec: Tensor, sc: SpecCfg) -> Tensor:                  # data_module.py:182-186
    if sc.spec_abs_exponent != 1:
        spec = spec.abs() ** sc.spec_abs_exponent * torch.exp(1j * spec.angle())
    return spec * sc.spec_factor


def spec_back(spec: Tensor, sc: SpecCfg) -> Tensor:                 # data_module.py:188-193
    spec = spec / sc.spec_factor
    if sc.spec_abs_exponent != 1:
        spec = spec.abs() ** (1 / sc.spec_abs_exponent) * torch.exp(1j * spec.angle())
    return spec


def pad_spec(Y: Tensor) -> Tensor:                                  # util/other.py:102-109
    T = Y.size(3)
    num_pad = (64 - T % 64) % 64
    return F.pad(Y, (0, num_pad, 0, 0))


def enhance(sd, cfg: NetCfg, y: Tensor, sde: OUVE = OUVE(), sc: SpecCfg = SpecCfg(), sampler_type: str = "pc",
            predictor: str = "reverse_diffusion", corrector: str = "ald", N: int = 50, corrector_steps: int = 1,
            snr: float = 0.5, t_eps: float = 3e-2, noise: NoiseFn = _default_noise, return_stft: bool = False):
    """ScoreModel.enhance (model.py:273-310) without the hard-wired .cuda(): y is a [1, L] waveform."""
    T_orig = y.size(1)
    norm_factor = y.abs().max().item()
    y = y / norm_factor
    Y = pad_spec(torch.unsqueeze(spec_fwd(stft(y, sc), sc), 0))
    score = make_score_fn(sd, cfg)
    if sampler_type == "pc":
        sample, nfe = pc_sample(score, Y, sde, N, predictor, corrector, corrector_steps, snr, t_eps, noise)
    elif sampler_type == "ode":
        sample, nfe = ode_sample(score, Y, sde, N, t_eps, noise=noise)
    else:
        raise ValueError(f"{sampler_type} is not a valid sampler type!")
    if return_stft:
        return sample.squeeze(), Y.squeeze(), T_orig, norm_factor
    x_hat = istft(spec_back(sample.squeeze(), sc), sc, T_orig) * norm_factor
    return x_hat.squeeze(), nfe


def enhance_storm(sd_den, sd_score, cfg_den: NetCfg, cfg_score: NetCfg, y: Tensor, sde: OUVE = OUVE(), sc: SpecCfg = SpecCfg(),
                  predictor: str = "reverse_diffusion", corrector: str = "none", N: int = 30, corrector_steps: int = 1, snr: float = 0.5,
                  t_eps: float = 3e-2, noise: NoiseFn = _default_noise, return_stft: bool = False):
    """StochasticRegenerationModel.enhance (model.py:720-780), condition="both": the denoiser estimate is the SDE mean
    and prior centre (:754), the score net sees cat[x, Y, Y_denoised] (:654-656,742-746)."""
    T_orig = y.size(1)
    norm_factor = y.abs().max().item()
    Y = pad_spec(torch.unsqueeze(spec_fwd(stft(y / norm_factor, sc), sc), 0))
    Y_den = ncsnpp_forward(sd_den, cfg_den, Y, None)

    def score(x, t, _y):
        return -ncsnpp_forward(sd_score, cfg_score, torch.cat([x, Y, Y_den], dim=1), t)

    sample, nfe = pc_sample(score, Y_den, sde, N, predictor, corrector, corrector_steps, snr, t_eps, noise)
    if return_stft:
        return sample.squeeze(), Y.squeeze(), Y_den.squeeze(), T_orig, norm_factor
    x_hat = istft(spec_back(sample.squeeze(), sc), sc, T_orig) * norm_factor
    return x_hat.squeeze(), nfe


def si_sdr(s, s_hat) -> float:                                      # util/other.py:82-86 (numpy form)
    s, s_hat = np.asarray(s, dtype=np.float64), np.asarray(s_hat, dtype=np.float64)
    alpha = np.dot(s_hat, s) / np.linalg.norm(s) ** 2
    return float(10 * np.log10(np.linalg.norm(alpha * s) ** 2 / np.linalg.norm(alpha * s - s_hat) ** 2))


def si_sdr_torch(s: Tensor, s_hat: Tensor) -> Tensor:               # util/other.py:88-94 (torch form, 1e-10 guards)
    n = min(s.size(-1), s_hat.size(-1))
    s, s_hat = s[..., :n], s_hat[..., :n]
    alpha = torch.dot(s_hat, s) / torch.norm(s) ** 2
    return 10 * torch.log10(1e-10 + torch.norm(alpha * s) ** 2 / (1e-10 + torch.norm(alpha * s - s_hat) ** 2))


def seeded_complex_noise(seed: int, count: int, shape) -> List[Tensor]:
    """``count`` draws of torch.randn_like(complex x) (real and imaginary parts N(0, 1/2)) from a seeded CPU generator:
    the deterministic noise source shared by the fixture generator (injected into the reference) and the tests."""
    g = torch.Generator().manual_seed(seed)
    return [torch.complex(torch.randn(*shape, generator=g), torch.randn(*shape, generator=g)) * math.sqrt(0.5) for _ in range(count)]


def synth_utterance(b: int, L: int) -> Tensor:
    """Synthetic noisy speech of SURVEY.md section 8(d): 5 harmonics of a random f0 with a 4 Hz on/off
    envelope plus white noise; generator seeded with 1234 + b.  Returns float32 [L]."""
    g = torch.Generator().manual_seed(1234 + b)
    f0 = 100.0 + 200.0 * torch.rand(1, generator=g).item()
    n = torch.arange(L, dtype=torch.float64) / 16000.0
    s = sum(torch.sin(2 * math.pi * f0 * (k + 1) * n) / (k + 1) for k in range(5))
    env = (torch.sin(2 * math.pi * 4.0 * n) > 0).double()
    noise = torch.randn(L, generator=g, dtype=torch.float64)
    return (0.1 * s * env + 0.05 * noise).float()
