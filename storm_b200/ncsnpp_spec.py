"""Host-side description of the NCSN++ score network: configuration, parameter table, seeded init.

The parameter table reproduces the *names, shapes and order* of the reference's ``state_dict``
(sgmse/backbones/ncsnpp.py:108-273 builds ``output_layer`` then ``all_modules``; block-internal order from
sgmse/backbones/ncsnpp_utils/layerspp.py:65-72,213-235) so reference checkpoints load unchanged and the
native engine can be fed one flat fp32 blob in exactly this order.
"""
from __future__ import annotations

from dataclasses import dataclass, asdict
from typing import Dict, List, Tuple

import numpy as np
import torch


@dataclass(frozen=True)
class NCSNppConfig:
    nf: int = 128
    ch_mult: Tuple[int, ...] = (1, 2, 2, 2)
    num_res_blocks: int = 1
    attn_resolutions: Tuple[int, ...] = (0,)
    image_size: int = 256
    input_channels: int = 4
    discriminative: bool = False
    fourier_scale: float = 16.0

    @property
    def conditional(self) -> bool:          # ncsnpp.py:81-86
        return not self.discriminative

    @property
    def in_ch(self) -> int:
        return 2 if self.discriminative else self.input_channels

    def to_dict(self):
        return asdict(self)


# options of the reference ctor (ncsnpp.py:40-65) whose non-default values select code this path does not build
_FIXED = dict(scale_by_sigma=True, nonlinearity="swish", resamp_with_conv=True, conditional=True, fir=True,
              fir_kernel=[1, 3, 3, 1], skip_rescale=True, resblock_type="biggan", progressive="output_skip",
              progressive_input="input_skip", progressive_combine="sum", embedding_type="fourier",
              spatial_channels=1, dropout=0.0, centered=False)


def config_from_kwargs(defaults: dict, kwargs: dict) -> NCSNppConfig:
    """Pick the topology options out of the kwargs soup the reference passes to every ctor
    (model.py:46-48); unknown keys are swallowed like ``**kwargs`` at ncsnpp.py:65."""
    kw = dict(defaults)
    for k in ("nf", "ch_mult", "num_res_blocks", "attn_resolutions", "image_size", "input_channels",
              "discriminative", "fourier_scale"):
        if k in kwargs and k not in defaults:
            kw[k] = kwargs[k]
    for k, v in _FIXED.items():
        if k in kwargs and kwargs[k] is not None:
            got = kwargs[k]
            got = list(got) if isinstance(got, tuple) else got
            if isinstance(got, str):
                got = got.lower()
            if got != v:
                raise NotImplementedError(
                    f"NCSN++ option {k}={kwargs[k]!r} is outside the B200 path (only {v!r} is built)")
    for k in ("ch_mult", "attn_resolutions"):
        if k in kw:
            kw[k] = tuple(int(v) for v in kw[k])
    return NCSNppConfig(**kw)


@dataclass(frozen=True)
class ModuleSpec:
    index: int            # position in all_modules
    kind: str             # fourier | linear | conv_in | res | attn | combine | gn | conv_out
    in_ch: int = 0
    out_ch: int = 0
    up: bool = False
    down: bool = False


def module_list(cfg: NCSNppConfig) -> List[ModuleSpec]:
    """The ``all_modules`` sequence for ``cfg`` (ncsnpp.py:110-273, default topology)."""
    nf, nres = cfg.nf, len(cfg.ch_mult)
    res = [cfg.image_size // (2 ** i) for i in range(nres)]
    mods: List[ModuleSpec] = []

    def add(kind, **kw):
        mods.append(ModuleSpec(len(mods), kind, **kw))

    add("fourier", out_ch=nf)
    if cfg.conditional:
        add("linear", in_ch=2 * nf, out_ch=4 * nf)
        add("linear", in_ch=4 * nf, out_ch=4 * nf)
    add("conv_in", in_ch=cfg.in_ch, out_ch=nf)
    hs_c = [nf]
    in_ch = nf
    for lvl in range(nres):
        for _ in range(cfg.num_res_blocks):
            out_ch = nf * cfg.ch_mult[lvl]
            add("res", in_ch=in_ch, out_ch=out_ch)
            in_ch = out_ch
            if res[lvl] in cfg.attn_resolutions:
                add("attn", in_ch=in_ch, out_ch=in_ch)
            hs_c.append(in_ch)
        if lvl != nres - 1:
            add("res", in_ch=in_ch, out_ch=in_ch, down=True)
            add("combine", in_ch=cfg.in_ch, out_ch=in_ch)
            hs_c.append(in_ch)
    in_ch = hs_c[-1]
    add("res", in_ch=in_ch, out_ch=in_ch)
    add("attn", in_ch=in_ch, out_ch=in_ch)
    add("res", in_ch=in_ch, out_ch=in_ch)
    for lvl in reversed(range(nres)):
        for _ in range(cfg.num_res_blocks + 1):
            out_ch = nf * cfg.ch_mult[lvl]
            add("res", in_ch=in_ch + hs_c.pop(), out_ch=out_ch)
            in_ch = out_ch
        if res[lvl] in cfg.attn_resolutions:
            add("attn", in_ch=in_ch, out_ch=in_ch)
        add("gn", in_ch=in_ch, out_ch=in_ch)
        add("conv_out", in_ch=in_ch, out_ch=cfg.in_ch)
        if lvl != 0:
            add("res", in_ch=in_ch, out_ch=in_ch, up=True)
    assert not hs_c
    return mods


def param_table(cfg: NCSNppConfig) -> List[Tuple[str, Tuple[int, ...]]]:
    """Ordered (key, shape) list == reference ``state_dict`` of NCSNpp(**cfg)."""
    out: List[Tuple[str, Tuple[int, ...]]] = [("output_layer.weight", (2, cfg.in_ch, 1, 1)), ("output_layer.bias", (2,))]
    temb_dim = 4 * cfg.nf
    for m in module_list(cfg):
        p = f"all_modules.{m.index}"
        if m.kind == "fourier":
            out.append((p + ".W", (m.out_ch,)))
        elif m.kind == "linear":
            out += [(p + ".weight", (m.out_ch, m.in_ch)), (p + ".bias", (m.out_ch,))]
        elif m.kind in ("conv_in", "conv_out"):
            out += [(p + ".weight", (m.out_ch, m.in_ch, 3, 3)), (p + ".bias", (m.out_ch,))]
        elif m.kind == "gn":
            out += [(p + ".weight", (m.in_ch,)), (p + ".bias", (m.in_ch,))]
        elif m.kind == "combine":
            out += [(p + ".Conv_0.weight", (m.out_ch, m.in_ch, 1, 1)), (p + ".Conv_0.bias", (m.out_ch,))]
        elif m.kind == "attn":
            out += [(p + ".GroupNorm_0.weight", (m.in_ch,)), (p + ".GroupNorm_0.bias", (m.in_ch,))]
            for i in range(4):
                out += [(p + f".NIN_{i}.W", (m.in_ch, m.in_ch)), (p + f".NIN_{i}.b", (m.in_ch,))]
        elif m.kind == "res":
            out += [(p + ".GroupNorm_0.weight", (m.in_ch,)), (p + ".GroupNorm_0.bias", (m.in_ch,)),
                    (p + ".Conv_0.weight", (m.out_ch, m.in_ch, 3, 3)), (p + ".Conv_0.bias", (m.out_ch,)),
                    (p + ".Dense_0.weight", (m.out_ch, temb_dim)), (p + ".Dense_0.bias", (m.out_ch,)),
                    (p + ".GroupNorm_1.weight", (m.out_ch,)), (p + ".GroupNorm_1.bias", (m.out_ch,)),
                    (p + ".Conv_1.weight", (m.out_ch, m.out_ch, 3, 3)), (p + ".Conv_1.bias", (m.out_ch,))]
            if m.in_ch != m.out_ch or m.up or m.down:
                out += [(p + ".Conv_2.weight", (m.out_ch, m.in_ch, 1, 1)), (p + ".Conv_2.bias", (m.out_ch,))]
        else:
            raise AssertionError(m.kind)
    return out


def param_count(cfg: NCSNppConfig) -> int:
    return sum(int(np.prod(s)) for _, s in param_table(cfg))


def random_state_dict(cfg: NCSNppConfig, seed: int = 0, init_scale: float = 1.0, bias_std: float = 0.0) -> Dict[str, torch.Tensor]:
    """Seeded random weights with the reference's init *distributions*
    (DDPM variance-scaling fan_avg uniform, layers.py:60-91; Fourier W ~ N(0, scale^2), layerspp.py:37),
    drawn from numpy's PCG64 so they are reproducible on any machine without the reference.

    ``init_scale`` plays the role of the reference's ``init_scale`` kwarg for Conv_1 / NIN_3 / pyramid
    output convs (default 0. there collapses the network to a constant - SURVEY.md section 0-5 - so the
    default here is 1.0).  ``bias_std`` > 0 additionally randomises every bias and GroupNorm affine so parity
    tests exercise those paths (the reference zero-initialises them).
    """
    rng = np.random.default_rng(seed)
    sd: Dict[str, torch.Tensor] = {}
    for key, shape in param_table(cfg):
        leaf = key.rsplit(".", 1)[-1]
        if key.endswith("all_modules.0.W"):
            v = rng.standard_normal(shape) * cfg.fourier_scale
        elif leaf in ("weight", "W") and len(shape) >= 2:
            if leaf == "W":                       # NIN: [in, out], default_init(scale=0.1) (layers.py:549-551)
                fan_in, fan_out, scale = shape[1], shape[0], 0.1      # reference quirk: in_axis=1,out_axis=0 on [in,out]
                if ".NIN_3." in key:
                    scale = init_scale if init_scale != 0 else 1e-10
            else:
                rf = int(np.prod(shape[2:])) if len(shape) > 2 else 1
                fan_in, fan_out, scale = shape[1] * rf, shape[0] * rf, 1.0
                if ".Conv_1." in key or _is_pyramid_out(key, cfg):
                    scale = init_scale if init_scale != 0 else 1e-10
            limit = np.sqrt(3.0 * scale / ((fan_in + fan_out) / 2.0))
            v = rng.uniform(-limit, limit, size=shape)
        elif leaf == "weight":                    # GroupNorm gamma
            v = 1.0 + bias_std * rng.standard_normal(shape)
        else:                                     # biases, GroupNorm beta, NIN b
            v = bias_std * rng.standard_normal(shape)
        sd[key] = torch.from_numpy(np.ascontiguousarray(v, dtype=np.float32))
    return sd


def _is_pyramid_out(key: str, cfg: NCSNppConfig) -> bool:
    idx = key.split(".")
    if idx[0] != "all_modules":
        return False
    i = int(idx[1])
    mods = module_list(cfg)
    return mods[i].kind == "conv_out"
