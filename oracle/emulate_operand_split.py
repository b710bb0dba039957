"""Emulate tensor-core operand splits in the CPU oracle, BEFORE writing a kernel for them.

TEST INFRASTRUCTURE ONLY (like everything under oracle/).  Every convolution / 1x1 / NIN product of the oracle network is
replaced by the sum of narrow-operand products a tensor-core kernel would issue (operands rounded with torch's float16 /
bfloat16 / float8 types, fp32 accumulation), the headline configuration (one utterance of BASELINE configs[1]: 256x256,
N=30 reverse_diffusion + ald = 60 evaluations, noise from ``seeded_complex_noise``) is run end to end and the enhanced
spectrogram is compared with the fixture produced by the UNMODIFIED reference (tests/golden/headline_256x256_N30.npz).

    python oracle/emulate_operand_split.py MODE        # a few minutes per mode on 8 cores

MODE                  tensor passes   what is multiplied                                               measured here
exact                 -               fp32 x fp32 (the oracle itself)                                  5e-7 .. 7e-7 (thread count)
f16_f8f8              2               wh16*h16 + e4m3(wl 2^12)*e5m2(x 2^-12) + e5m2(w 2^-6)*e4m3(l 2^6)   5.8e-5   <- conv_tc.cu mode 1
f16_f8f8_noscale_l    2               the same with e4m3(w) * e5m2(l): no scaling of the residual (one multiply less per element)   6.9e-5
skip_f16_act          2 (skip: 1.5)   f16_f8f8, but the raw 1x1 skip convs drop the residual plane l    1.9e-4
f16_act               2               (wh16 + wl16) * h16: fp16 activations, exact weights             8.0e-4
f16_main              1               wh16 * h16 only                                                  1.6e-3
f16_mxf4              1.5 (if kind::mxf4 runs at twice the fp8 rate)   wh16*h16 + mxfp4(wl)*mxfp4(h) + mxfp4(w)*mxfp4(l): both correction
                                      products in OCP MXFP4 (e2m1, one power-of-two scale per 32 channels)   1.8e-4   
f16_nvf4              1.5             the same in NVFP4 (e2m1, one e4m3 scale per 16 channels + a per-tensor power of two)   1.8e-4
f16_mxf4_fixed        1.5             activation fp4 planes with ONE power-of-two scale per sample (from its rms), weights in 32-channel blocks   4.5e-4
f16_mxf4_pix_tied     1.5             f16_mxf4_pix with the residual plane's exponent tied to the main plane's (exponent - 11): one factor per pixel   2.4e-4
f16_mxf4_pix          1.5             f16_mxf4 with ONE exponent per pixel for all channels of the activation planes (weights: 32-channel blocks)   2.2e-4
(bf16 x 3, conv_tc.cu mode 0, is measured on the GPU: 2.8e-5.)  The north-star tolerance is 1e-3.
"""
import os
import sys
import time

import numpy as np
import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import storm_oracle as O  # noqa: E402
from storm_b200 import ncsnpp_spec  # noqa: E402

E4, E5 = torch.float8_e4m3fn, torch.float8_e5m2


def q8(x, dt, k):
    """x * 2^k rounded to fp8 (saturating, round to nearest even), returned unscaled."""
    s = 2.0 ** k
    lim = 440.0 if dt is E4 else 57000.0
    return (x * s).clamp(-lim, lim).to(dt).float() / s


_E2M1 = torch.tensor([0.0, 0.5, 1.0, 1.5, 2.0, 3.0, 4.0, 6.0])


def q4(x, dim, nv, whole=False):
    """Block-scaled fp4 (e2m1) along `dim` (the K dimension = input channels): OCP MXFP4 (blocks of 32, power-of-two scale chosen
    as floor(log2(max)) - 2, values above 6 saturate) or NVFP4 (blocks of 16, scale = e4m3(max / 6))."""
    xm = x.movedim(dim, -1)
    shp = xm.shape
    C = shp[-1]
    blk = C if whole else (16 if nv else 32)          # whole: ONE scale per row (e.g. per pixel for all channels)
    pad = (-C) % blk
    if pad:
        xm = F.pad(xm, (0, pad))
    xb = xm.reshape(*xm.shape[:-1], -1, blk)
    amax = xb.abs().amax(dim=-1, keepdim=True).clamp_min(1e-30)
    if nv:
        # NVFP4's e4m3 block scales only span 2^-9 .. 448: a per-tensor power of two (the format's second-level scale) centres them
        g = torch.exp2(torch.floor(torch.log2(amax.max())) - 6.0)
        scale = (amax / (6.0 * g)).clamp(2.0 ** -9, 448.0).to(E4).float().clamp_min(2.0 ** -9) * g
    else:
        scale = torch.exp2(torch.floor(torch.log2(amax)) - 2.0)
    y = (xb / scale).clamp(-6.0, 6.0)
    mag = y.abs()
    idx = torch.bucketize(mag, (_E2M1[:-1] + _E2M1[1:]) / 2)          # round to nearest representable magnitude
    q = _E2M1[idx] * torch.sign(y) * scale
    q = q.reshape(*xm.shape)[..., :C].reshape(shp)
    return q.movedim(-1, dim)


def make_lin(mode):
    def lin(op, x, w):
        if mode == "exact":
            return op(x, w)
        h = x.half().float()
        wh = w.half().float()
        if mode == "f16_act":
            return op(h, w)
        if mode == "f16_main":
            return op(h, wh)
        if mode == "f16_f8f8":
            return op(h, wh) + op(q8(x, E5, -12), q8(w - wh, E4, 12)) + op(q8(x - h, E4, 6), q8(w, E5, -6))
        if mode == "f16_f8f8_noscale_l":      # residual unscaled in e5m2, weights unscaled in e4m3 (one multiply less per element)
            return op(h, wh) + op(q8(x, E5, -12), q8(w - wh, E4, 12)) + op(q8(x - h, E5, 0), q8(w, E4, 0))
        if mode == "f16_mxf4_fixed":
            # no activation exponents at all: ONE power of two per sample (from the tensor's rms, which GroupNorm already measures),
            # values above 6 * scale saturate; residual plane at scale * 2^-11
            kd = 0 if w.dim() == 2 else 1
            rms = h.pow(2).mean(dim=(1, 2, 3), keepdim=True).sqrt().clamp_min(1e-30)
            sc = torch.exp2(torch.round(torch.log2(rms)))

            def q4f(v, s_):
                y = (v / s_).clamp(-6.0, 6.0)
                idx = torch.bucketize(y.abs().contiguous(), (_E2M1[:-1] + _E2M1[1:]) / 2)
                return _E2M1[idx] * torch.sign(y) * s_
            return op(h, wh) + op(q4f(h, sc), q4(w - wh, kd, False)) + op(q4f(x - h, sc * 2.0 ** -11), q4(w, kd, False))
        if mode == "f16_mxf4_pix_tied":
            # the kernel-shaped variant: ONE exponent per pixel from max|h| over all channels; the residual plane uses the same
            # exponent - 11 (|l| <= 2^-11 |h| element-wise), so a single per-pixel factor multiplies the whole correction accumulator
            kd = 0 if w.dim() == 2 else 1
            amax = h.abs().amax(dim=1, keepdim=True).clamp_min(1e-30)
            sc = torch.exp2(torch.floor(torch.log2(amax)) - 2.0)

            def q4s(v, s_):
                y = (v / s_).clamp(-6.0, 6.0)
                idx = torch.bucketize(y.abs().contiguous(), (_E2M1[:-1] + _E2M1[1:]) / 2)
                return _E2M1[idx] * torch.sign(y) * s_
            return op(h, wh) + op(q4s(h, sc), q4(w - wh, kd, False)) + op(q4s(x - h, sc * 2.0 ** -11), q4(w, kd, False))
        if mode in ("f16_mxf4", "f16_nvf4", "f16_mxf4_pix"):
            nv = mode == "f16_nvf4"
            pix = mode == "f16_mxf4_pix"          # activations: one exponent per pixel for all channels (weights keep 32-channel blocks)
            kd = 0 if w.dim() == 2 else 1          # NIN weights are [Cin, Cout], conv weights [Cout, Cin, kh, kw]
            return op(h, wh) + op(q4(h, 1, nv, pix), q4(w - wh, kd, nv)) + op(q4(x - h, 1, nv, pix), q4(w, kd, nv))
        raise SystemExit(f"unknown mode {mode}")
    return lin


def main():
    mode = sys.argv[1] if len(sys.argv) > 1 else "f16_f8f8"
    torch.set_num_threads(min(16, os.cpu_count() or 1))
    lin = make_lin("f16_f8f8" if mode == "skip_f16_act" else mode)
    lin_skip = make_lin("f16_act") if mode == "skip_f16_act" else lin      # 1x1 skip convs (Conv_2) without a residual plane

    def conv(x, sd, key, pad):
        f = lin_skip if key.endswith(".Conv_2") else lin
        return f(lambda a, b: F.conv2d(a, b, None, padding=pad), x, sd[key + ".weight"]) + sd[key + ".bias"][None, :, None, None]

    O._conv = conv
    O._nin = lambda x, sd, key: lin(lambda a, b: torch.einsum("bchw,cd->bdhw", a, b), x, sd[key + ".W"]) \
        + sd[key + ".b"][None, :, None, None]
    g = np.load(os.path.join(ROOT, "tests", "golden", "headline_256x256_N30.npz"))
    sd = ncsnpp_spec.random_state_dict(ncsnpp_spec.NCSNppConfig(), seed=int(g["weight_seed"]), init_scale=float(g["init_scale"]),
                                       bias_std=float(g["bias_std"]))
    sc = O.SpecCfg()
    yw = O.synth_utterance(0, int(g["L"])).reshape(1, -1)
    Y = O.pad_spec(O.spec_fwd(O.stft(yw / yw.abs().max(), sc), sc).unsqueeze(0))
    N = int(g["N"])
    it = iter(O.seeded_complex_noise(int(g["noise_seed"]), 1 + 2 * N, tuple(Y.shape)))
    t0 = time.time()
    with torch.no_grad():
        x, nfe = O.pc_sample(O.make_score_fn(sd, O.NetCfg()), Y, O.OUVE(), N, noise=lambda z: next(it))
    ref = torch.from_numpy(g["sample"])
    print(f"mode={mode}: {nfe} evaluations, enhanced spectrogram rel-L2 vs the reference = {(x - ref).norm() / ref.norm():.3e}"
          f"  ({time.time() - t0:.0f} s)")


if __name__ == "__main__":
    main()
