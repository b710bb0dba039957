"""Backbone plugin boundary: ``BackboneRegistry`` and the NCSN++ score networks behind it.

Mirrors sgmse/backbones/shared.py:17 (the registry) and sgmse/backbones/ncsnpp.py:36-513 (the registered
classes and their ctor / forward contract).  The classes are ``torch.nn.Module``s only so that
``state_dict`` / ``load_state_dict`` / ``parameters`` / ``.cuda()`` keep working with reference checkpoints
(same keys, 271 entries for ``ncsnpp``); ``forward`` hands the tensors to the native sm_100a engine through
the C ABI - there is no PyTorch implementation of the network in this package.
"""
from __future__ import annotations

from typing import Optional

import torch
from torch import nn

from . import _native, ncsnpp_spec
from .registry import Registry

BackboneRegistry = Registry("Backbone")


def _ensure_path(root: nn.Module, parts):
    mod = root
    for part in parts:
        if part not in mod._modules:
            mod.add_module(part, nn.ModuleList() if part == "all_modules" else nn.Module())
        mod = mod._modules[part]
    return mod


@BackboneRegistry.register("ncsnpp")
class NCSNpp(nn.Module):
    """NCSN++ (27.8 M parameters with the defaults).  kwargs follow ncsnpp.py:40-65; unknown ones are
    swallowed, options that select code outside this path raise NotImplementedError."""

    DEFAULTS: dict = {}

    def __init__(self, seed: int = 0, init_scale: float = 0.0, engine: int = 0, **kwargs):
        super().__init__()
        self.FORCE_STFT_OUT = False                     # read by the reference's DiscriminativeModel (model.py:323)
        self.cfg = ncsnpp_spec.config_from_kwargs(dict(self.DEFAULTS), kwargs)
        self.discriminative = self.cfg.discriminative
        self._engine_kind = engine
        self._engine: Optional[_native.Engine] = None
        self._dirty = True
        # parameters under the reference's names; default init reproduces the reference distributions
        # (init_scale=0. -> 1e-10 variance scale for Conv_1 / NIN_3 / pyramid convs, ncsnpp.py:56, layers.py:90)
        for key, value in ncsnpp_spec.random_state_dict(self.cfg, seed=seed, init_scale=init_scale).items():
            *path, leaf = key.split(".")
            _ensure_path(self, path).register_parameter(leaf, nn.Parameter(value, requires_grad=False))

    @staticmethod
    def add_argparse_args(parser):
        return parser

    # ---- weights -> native engine ------------------------------------------------------------------------
    def load_state_dict(self, state_dict, strict: bool = True, **kw):
        out = super().load_state_dict(state_dict, strict=strict, **kw)
        self._dirty = True
        return out

    def _apply(self, fn, *a, **kw):
        out = super()._apply(fn, *a, **kw)
        self._dirty = True
        return out

    def mark_weights_changed(self):
        self._dirty = True

    def flat_weights(self, table, total: int) -> torch.Tensor:
        """The engine's blob: tensor ``name`` at its (64-float aligned) offset, zeros in the gaps."""
        sd = self.state_dict()
        dev = next(iter(sd.values())).device
        blob = torch.zeros(total, dtype=torch.float32, device=dev)
        for name, shape, off in table:
            v = sd[name]
            blob[off:off + v.numel()] = v.reshape(-1).to(torch.float32)
        return blob

    def engine(self) -> _native.Engine:
        dev = next(self.parameters()).device
        if dev.type != "cuda":
            raise RuntimeError("storm_b200: move the backbone to a CUDA device first (.cuda()); the path has no CPU implementation")
        if self._engine is not None and self._engine.device != dev:      # .cuda(other) / .to(other): the handle follows the weights
            self._engine.close()
            self._engine = None
        if self._engine is None:
            c = self.cfg
            with torch.cuda.device(dev):
                self._engine = _native.Engine(nf=c.nf, ch_mult=c.ch_mult, num_res_blocks=c.num_res_blocks,
                                              attn_resolutions=c.attn_resolutions, image_size=c.image_size,
                                              input_channels=c.input_channels, discriminative=c.discriminative,
                                              engine=self._engine_kind)
            self._table = self._engine.param_table()
            if [(n, s) for n, s, _ in self._table] != ncsnpp_spec.param_table(c):
                raise RuntimeError("native parameter table disagrees with the Python specification")
            self._dirty = True
        if self._dirty:
            self._engine.load_blob(self.flat_weights(self._table, self._engine.param_floats()))
            self._dirty = False
        return self._engine

    def forward(self, x: torch.Tensor, time_cond: Optional[torch.Tensor] = None) -> torch.Tensor:
        """x: complex64 [B, input_channels/2, F, T] (x and y stacked on dim 1, model.py:127-129);
        time_cond: float [B].  Returns complex64 [B, 1, F, T] (ncsnpp.py:281-284,446-450)."""
        return self.engine().net_forward(x, time_cond)


@BackboneRegistry.register("ncsnpplarge", "ncsnpp-large")     # README.md:43 of the reference spells it "ncsnpp-large"
class NCSNppLarge(NCSNpp):
    """65.6 M parameters (ncsnpp.py:460-470)."""
    DEFAULTS = dict(nf=128, ch_mult=(1, 1, 2, 2, 2, 2, 2), num_res_blocks=2, attn_resolutions=(16,))




@BackboneRegistry.register("ncsnpp12M")
class NCSNpp12M(NCSNpp):
    """ncsnpp.py:479-489"""
    DEFAULTS = dict(nf=96, ch_mult=(1, 2, 2, 1), num_res_blocks=1, attn_resolutions=(0,))


@BackboneRegistry.register("ncsnpp6M")
class NCSNpp6M(NCSNpp):
    """ncsnpp.py:498-508"""
    DEFAULTS = dict(nf=96, ch_mult=(1, 1, 1, 1), num_res_blocks=1, attn_resolutions=(0,))
