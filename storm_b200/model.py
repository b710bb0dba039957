"""``ScoreModel``: the reference-facing entry point of the path (sgmse/model.py:24-310).

Same constructor keywords, ``forward`` / ``get_pc_sampler`` / ``get_ode_sampler`` / ``enhance`` signatures and
return forms; ``load_from_checkpoint`` reads the reference's Lightning ``.ckpt`` files (state_dict under
``dnn.*`` plus the ``torch_ema`` shadow parameters, model.py:86-111) without needing pytorch_lightning.
Training code (losses, optimisers, dataloaders) is out of scope.  New: ``enhance_batch`` (the reference
handles one utterance per call) and sharding helpers used by ``enhancement.py --gpus``.
"""
from __future__ import annotations

import time
from math import ceil
from typing import List, Optional, Sequence, Union

import torch

from . import _native, sampling
from .backbones import BackboneRegistry
from .data_module import SpecsDataModule
from .sdes import SDERegistry


FROZEN_PARAMS = ("all_modules.0.W",)      # GaussianFourierProjection.W: requires_grad=False (layerspp.py:37)


def map_ema_shadow(named_params, shadow):
    """Pair ``ema.shadow_params`` of a reference checkpoint with the parameters they belong to.

    ``torch_ema`` keeps its shadow list positionally, in ``model.parameters()`` order (sgmse/model.py:55: the EMA is built
    from ``self.parameters()``; tests/golden/param_order.json pins that order).  Releases differ on whether frozen
    parameters get a shadow: 0.2 filters on ``requires_grad``, 0.3 (requirements.txt:15) shadows every parameter.  The
    only frozen parameter on this path is the Fourier ``W``.  Both layouts are accepted, every pair is shape-checked, and
    anything else raises instead of silently mis-assigning weights.  Returns ``[(name, shadow_tensor), ...]``."""
    named_params = list(named_params)
    shadow = list(shadow)
    trainable = [(n, p) for n, p in named_params if n.split("dnn.")[-1].split("_net.")[-1] not in FROZEN_PARAMS]
    if len(shadow) == len(named_params):
        target = named_params
    elif len(shadow) == len(trainable):
        target = trainable
    else:
        raise RuntimeError(f"checkpoint EMA has {len(shadow)} shadow parameters; the model has {len(named_params)} parameters "
                           f"({len(trainable)} trainable) - cannot align them")
    out = []
    for (n, p), sh in zip(target, shadow):
        if tuple(sh.shape) != tuple(p.shape):
            raise RuntimeError(f"checkpoint EMA shadow for '{n}' has shape {tuple(sh.shape)}, expected {tuple(p.shape)}")
        out.append((n, sh))
    return out


class ScoreModel:
    def __init__(self, backbone: str = "ncsnpp", sde: str = "ouve", lr: float = 1e-4, ema_decay: float = 0.999,
                 t_eps: float = 3e-2, transform: str = "none", nolog: bool = False, num_eval_files: int = 50,
                 loss_type: str = "mse", data_module_cls=None, seed: int = 42, **kwargs):
        if sde == "ouvesde":          # the reference's own (unregistered) default spelling, model.py:26
            sde = "ouve"
        dnn_cls = BackboneRegistry.get_by_name(backbone)
        kwargs.update(input_channels=4)                       # model.py:47
        self.dnn = dnn_cls(**kwargs)
        self.sde = SDERegistry.get_by_name(sde)(**kwargs)
        self.t_eps = t_eps
        self.data_module = (data_module_cls or SpecsDataModule)(**kwargs)
        self.hparams = dict(backbone=backbone, sde=sde, t_eps=t_eps, **kwargs)
        self._seed = seed
        self._calls = 0
        self._ema_shadow: Optional[list] = None          # [(parameter name, shadow tensor)] from map_ema_shadow
        self._raw_params: Optional[dict] = None
        self.device = torch.device("cpu")

    # ---- nn.Module-like plumbing ----------------------------------------------------------------------------
    def cuda(self, device=None):
        self.dnn.cuda(device)
        self.device = next(self.dnn.parameters()).device
        return self

    def to(self, *args, **kwargs):
        self.dnn.to(*args, **kwargs)
        self.device = next(self.dnn.parameters()).device
        return self

    def parameters(self):
        return self.dnn.parameters()

    def state_dict(self):
        return {"dnn." + k: v for k, v in self.dnn.state_dict().items()}

    def load_state_dict(self, sd, strict=True):
        return self.dnn.load_state_dict({k[4:]: v for k, v in sd.items() if k.startswith("dnn.")}, strict=strict)

    def eval(self, no_ema: bool = False):
        """Swap the EMA shadow weights in for inference unless ``no_ema`` (model.py:97-111)."""
        if self._ema_shadow is not None and not no_ema and self._raw_params is None:
            params = dict(self.dnn.named_parameters())
            self._raw_params = {n: params[n].detach().clone() for n, _ in self._ema_shadow}
            with torch.no_grad():
                for n, s in self._ema_shadow:
                    params[n].copy_(s.to(params[n].device))
            self.dnn.mark_weights_changed()
        return self

    def train(self, mode: bool = True, no_ema: bool = False):
        if mode:
            if self._raw_params is not None:
                params = dict(self.dnn.named_parameters())
                with torch.no_grad():
                    for n, s in self._raw_params.items():
                        params[n].copy_(s)
                self._raw_params = None
                self.dnn.mark_weights_changed()
            return self
        return self.eval(no_ema=no_ema)

    @classmethod
    def load_from_checkpoint(cls, checkpoint_path, map_location="cpu", **overrides):
        """Read a reference Lightning checkpoint (enhancement.py:56-60)."""
        ckpt = _load_ckpt(checkpoint_path, map_location)
        hp = dict(ckpt.get("hyper_parameters", {}))
        hp.pop("data_module_cls", None)
        hp.update({k: v for k, v in overrides.items() if k not in ("kwargs",)})
        model = cls(**hp)
        model.load_state_dict(ckpt["state_dict"], strict=True)
        ema = ckpt.get("ema")
        if ema is not None and ema.get("shadow_params") is not None:
            model._ema_shadow = [(n, s.detach().clone()) for n, s in map_ema_shadow(model.dnn.named_parameters(), ema["shadow_params"])]
        return model

    def next_seed(self) -> int:
        self._calls += 1
        return (self._seed * 1000003 + self._calls) & 0xFFFFFFFFFFFF

    # ---- score ------------------------------------------------------------------------------------------------
    def _raw_dnn_output(self, x, t, y):
        return self.dnn(torch.cat([x, y], dim=1), t)                 # model.py:127-129

    def forward(self, x, t, y, **kwargs):
        return -self._raw_dnn_output(x, t, y)                        # model.py:131-136

    __call__ = forward

    # ---- samplers ---------------------------------------------------------------------------------------------
    def get_pc_sampler(self, predictor_name, corrector_name, y, N=None, minibatch=None, scale_factor=None, **kwargs):
        sde = self.sde.copy()
        sde.N = self.sde.N if N is None else N
        kwargs = {"eps": self.t_eps, **kwargs}
        if minibatch is None:
            return sampling.get_pc_sampler(predictor_name, corrector_name, sde=sde, score_fn=self, y=y, **kwargs)
        return self._minibatched(lambda ym, ids: sampling.get_pc_sampler(predictor_name, corrector_name, sde=sde, score_fn=self,
                                                                        y=ym, utt_ids=ids, **kwargs), y, minibatch)

    def get_ode_sampler(self, y, N=None, minibatch=1, **kwargs):
        sde = self.sde.copy()
        sde.N = self.sde.N if N is None else N
        kwargs = {"eps": self.t_eps, **kwargs}
        if minibatch is None:
            return sampling.get_ode_sampler(sde, self, y=y, **kwargs)
        return self._minibatched(lambda ym, ids: sampling.get_ode_sampler(sde, self, y=ym, utt_ids=ids, **kwargs), y, minibatch)

    @staticmethod
    def _minibatched(make, y, minibatch):                            # model.py:208-222
        M = y.shape[0]

        def batched_sampling_fn():
            samples, ns = [], []
            for i in range(int(ceil(M / minibatch))):
                lo, hi = i * minibatch, min(M, (i + 1) * minibatch)
                sample, n = make(y[lo:hi], list(range(lo, hi)))()
                samples.append(sample)
                ns.append(n)
            return torch.cat(samples, dim=0), ns

        return batched_sampling_fn

    # ---- audio <-> spectrogram ----------------------------------------------------------------------------------
    def _forward_transform(self, spec):
        return self.data_module.spec_fwd(spec)

    def _backward_transform(self, spec):
        return self.data_module.spec_back(spec)

    def _stft(self, sig):
        return self.data_module.stft(sig)

    def _istft(self, spec, length=None):
        return self.data_module.istft(spec, length)

    def to_audio(self, spec, length=None):
        return self._istft(self._backward_transform(spec), length)

    # ---- one-call enhancement ---------------------------------------------------------------------------------
    def enhance(self, y, sampler_type="pc", predictor="reverse_diffusion", corrector="ald", N=50, corrector_steps=1, snr=0.5,
                timeit=False, scale_factor=None, return_stft=False, **kwargs):
        """One utterance ``y`` [1, L] -> enhanced waveform [L] on the CPU, like model.py:273-310."""
        start = time.time()
        T_orig = y.size(1)
        dev = self._device()
        yd = y.to(dev, torch.float32)
        if sampler_type == "pc" and not return_stft and not kwargs:
            sc = self._sampler_cfg(predictor, corrector, N, corrector_steps, snr)
            x_hat, nfe = self.dnn.engine().enhance(yd, sc, self.data_module.spec_config(), seed=self.next_seed())
            x_hat = x_hat.squeeze().cpu()
        else:
            norm = _native.wave_norm(yd)
            T_pad = _native.padded_frames(_native.num_frames(T_orig, self.data_module.spec_config()))
            Y = _native.stft(yd, norm, self.data_module.spec_config(), T_pad).unsqueeze(1)
            if sampler_type == "pc":
                sampler = self.get_pc_sampler(predictor, corrector, Y, N=N, corrector_steps=corrector_steps, snr=snr, **kwargs)
            elif sampler_type == "ode":
                sampler = self.get_ode_sampler(Y, N=N, **kwargs)
            else:
                raise ValueError("{} is not a valid sampler type!".format(sampler_type))
            sample, nfe = sampler()
            if return_stft:
                return sample.squeeze(), Y.squeeze(), T_orig, norm.item()
            x_hat = self.dnn.engine().istft(sample.squeeze(1), norm, T_orig, self.data_module.spec_config()).squeeze().cpu()
        if timeit:
            return x_hat, nfe, (time.time() - start) / (len(x_hat) / 16000)
        return x_hat

    def enhance_batch(self, ys: Union[torch.Tensor, Sequence[torch.Tensor]], predictor="reverse_diffusion", corrector="ald", N=50,
                      corrector_steps=1, snr=0.5, seed: Optional[int] = None, utt_ids: Optional[Sequence[int]] = None,
                      return_nfe: bool = False, max_batch: int = 8):
        """Batched ``enhance`` (the reference has none: enhancement.py:66-72 loops over files).

        * ``ys`` = tensor [B, L] (or a list of equal-length [L] / [1, L]) -> tensor [B, L] on the input's device;
        * ``ys`` = list of utterances of DIFFERENT lengths -> list of [L_b] tensors: utterances are bucketed by padded
          frame count (``parallel.bucket_utterances``), each bucket runs as ragged batches of at most ``max_batch``
          through ``storm_enhance_ragged``; every utterance gets exactly the frames, padding and istft length the
          reference gives it when enhanced alone.
        Host tensors are staged through pinned memory; noise is keyed per utterance id, so results do not depend on
        how a set of utterances is split into batches or across GPUs."""
        sc = self._sampler_cfg(predictor, corrector, N, corrector_steps, snr)
        sp = self.data_module.spec_config()
        seed = self.next_seed() if seed is None else seed
        dev = self._device()
        if not torch.is_tensor(ys):
            ys = [y.reshape(-1) for y in ys]
            lens = [int(y.numel()) for y in ys]
            if len(set(lens)) > 1:
                from .parallel import bucket_utterances
                ids_all = list(range(len(ys))) if utt_ids is None else [int(u) for u in utt_ids]
                outs: List[Optional[torch.Tensor]] = [None] * len(ys)
                nfe = 0
                for idx in bucket_utterances(lens, sp.hop_length, max_batch):
                    Lmax = max(lens[i] for i in idx)
                    host = torch.zeros(len(idx), Lmax, dtype=torch.float32)
                    for r, i in enumerate(idx):
                        host[r, :lens[i]] = ys[i].to("cpu", torch.float32)
                    yd = host.pin_memory().to(dev, non_blocking=True)
                    out, nfe = self.dnn.engine().enhance(yd, sc, sp, seed=seed, utt_ids=[ids_all[i] for i in idx],
                                                         lengths=[lens[i] for i in idx])
                    for r, i in enumerate(idx):
                        outs[i] = out[r, :lens[i]].to(ys[i].device)
                return (outs, nfe) if return_nfe else outs
            ys = torch.stack(ys)
        src_dev = ys.device
        yd = ys.to(torch.float32)
        if not yd.is_cuda:
            yd = yd.pin_memory().to(dev, non_blocking=True)
        out, nfe = self.dnn.engine().enhance(yd, sc, sp, seed=seed, utt_ids=utt_ids)
        out = out.to(src_dev)
        return (out, nfe) if return_nfe else out

    def _sampler_cfg(self, predictor, corrector, N, corrector_steps, snr):
        s = self.sde
        return _native.make_sampler_config(predictor, corrector, N, corrector_steps, snr, self.t_eps, s.theta, s.sigma_min, s.sigma_max)

    def _device(self):
        p = next(self.dnn.parameters())
        if not p.is_cuda:
            raise RuntimeError("ScoreModel.enhance needs the model on a CUDA device: call model.cuda() first "
                               "(the reference hard-codes y.cuda(), model.py:285)")
        return p.device


class DiscriminativeModel(ScoreModel):
    """Predictive baseline of the reference (sgmse/model.py:320-370): the same backbone called once with ``t = 1``, no
    sampler.  Checkpoints are ScoreModel-shaped (``dnn.*`` keys, ``backbone`` hyper-parameter) and are what
    ``enhancement.py --mode denoiser-only`` loads (enhancement.py:54-55)."""

    def __init__(self, *args, **kwargs):
        kwargs.setdefault("discriminative", False)     # the stock DiscriminativeModel keeps the 4-channel conditional NCSN++
        super().__init__(*args, **kwargs)

    def forward(self, y):                                            # model.py:322-327 (FORCE_STFT_OUT is False for NCSN++)
        t = torch.ones(y.shape[0], device=y.device)
        return self.dnn(y, t)

    __call__ = forward

    def _net_input(self, Y):
        """ScoreModel.__init__ forces ``input_channels=4`` (model.py:47) while ``forward`` hands the backbone Y alone
        (model.py:326): with the stock 4-channel NCSN++ the reference's own call fails on the channel count, so published
        discriminative checkpoints necessarily carry a backbone whose real-valued input has 2 channels
        (``discriminative=True``).  Anything else is reported instead of guessed."""
        if self.dnn.cfg.in_ch != 2 * Y.shape[1]:
            raise RuntimeError(f"DiscriminativeModel: the backbone expects {self.dnn.cfg.in_ch} real input channels but the noisy "
                               f"spectrogram provides {2 * Y.shape[1]} (the reference's DiscriminativeModel.forward has the same "
                               "mismatch with a 4-channel backbone, sgmse/model.py:326)")
        return Y

    def enhance(self, y, **ignored_kwargs):
        """model.py:351-370 for one utterance ``y`` [1, L] -> [L] on the CPU."""
        return self.enhance_batch(y.reshape(1, -1))[0].cpu()

    def enhance_batch(self, ys, **ignored_kwargs):
        """[B, L] (or list of equal-length utterances) -> [B, L]: normalise, STFT, one network call, iSTFT."""
        if not torch.is_tensor(ys):
            ys = torch.stack([y.reshape(-1) for y in ys])
        dev = self._device()
        src = ys.device
        yd = ys.to(torch.float32)
        yd = yd.pin_memory().to(dev, non_blocking=True) if not yd.is_cuda else yd
        sp = self.data_module.spec_config()
        norm = _native.wave_norm(yd)
        T_pad = _native.padded_frames(_native.num_frames(yd.shape[1], sp))
        Y = _native.stft(yd, norm, sp, T_pad).unsqueeze(1)
        X = self.forward(self._net_input(Y))
        return self.dnn.engine().istft(X.squeeze(1), norm, yd.shape[1], sp).to(src)


class StochasticRegenerationModel:
    """StoRM inference (sgmse/model.py:392-780): a discriminative NCSN++ first estimates ``Y_denoised``; the score
    network, conditioned on ``[Y, Y_denoised]`` (6 real input channels), then runs reverse diffusion around
    ``Y_denoised``.  Same ctor keywords / ``enhance`` signature and defaults (corrector="none", N=30) as the
    reference; training parts are out of scope."""

    def __init__(self, backbone_denoiser: str = "ncsnpp", backbone_score: str = "ncsnpp", sde: str = "ouve", lr: float = 1e-4,
                 ema_decay: float = 0.999, t_eps: float = 3e-2, nolog: bool = False, num_eval_files: int = 50,
                 loss_type_denoiser: str = "none", loss_type_score: str = "mse", data_module_cls=None, mode="regen-joint-training",
                 condition="both", seed: int = 42, **kwargs):
        if condition not in ("noisy", "post_denoiser", "both"):
            raise NotImplementedError(f"Don't know the conditioning you have wished for: {condition}")
        kw_d = dict(kwargs, input_channels=2, discriminative=True)                    # model.py:416-419
        self.denoiser_net = BackboneRegistry.get_by_name(backbone_denoiser)(**kw_d) if backbone_denoiser != "none" else None
        kw_s = dict(kwargs, input_channels=(6 if condition == "both" else 4), discriminative=False)   # model.py:421-423
        self.score_net = BackboneRegistry.get_by_name(backbone_score)(**kw_s) if backbone_score != "none" else None
        self.sde = SDERegistry.get_by_name("ouve" if sde == "ouvesde" else sde)(**kwargs)
        self.t_eps, self.condition, self.mode = t_eps, condition, mode
        self.data_module = (data_module_cls or SpecsDataModule)(**kwargs)
        self._seed, self._calls = seed, 0
        self._ema_shadow: Optional[list] = None
        self._raw_params: Optional[dict] = None

    def named_parameters(self):
        """``self.parameters()`` order of the reference module: denoiser_net then score_net (model.py:416-423)."""
        for name, net in (("denoiser_net", self.denoiser_net), ("score_net", self.score_net)):
            if net is not None:
                for n, p in net.named_parameters():
                    yield f"{name}.{n}", p

    def _nets(self):
        return [n for n in (self.denoiser_net, self.score_net) if n is not None]

    def cuda(self, device=None):
        for n in self._nets():
            n.cuda(device)
        return self

    def eval(self, no_ema: bool = False):
        """Swap the EMA shadow weights in unless ``no_ema`` (model.py:517-531)."""
        if self._ema_shadow is not None and not no_ema and self._raw_params is None:
            params = dict(self.named_parameters())
            self._raw_params = {n: params[n].detach().clone() for n, _ in self._ema_shadow}
            with torch.no_grad():
                for n, sh in self._ema_shadow:
                    params[n].copy_(sh.to(params[n].device))
            for n in self._nets():
                n.mark_weights_changed()
        return self

    def train(self, mode: bool = True, no_ema: bool = False):
        if not mode:
            return self.eval(no_ema=no_ema)
        if self._raw_params is not None:
            params = dict(self.named_parameters())
            with torch.no_grad():
                for n, v in self._raw_params.items():
                    params[n].copy_(v)
            self._raw_params = None
            for n in self._nets():
                n.mark_weights_changed()
        return self

    def state_dict(self):
        out = {}
        for name, net in (("denoiser_net", self.denoiser_net), ("score_net", self.score_net)):
            if net is not None:
                out.update({f"{name}.{k}": v for k, v in net.state_dict().items()})
        return out

    def load_state_dict(self, sd, strict=True):
        for name, net in (("denoiser_net", self.denoiser_net), ("score_net", self.score_net)):
            if net is not None:
                net.load_state_dict({k[len(name) + 1:]: v for k, v in sd.items() if k.startswith(name + ".")}, strict=strict)

    @classmethod
    def load_from_checkpoint(cls, checkpoint_path, map_location="cpu", **overrides):
        ckpt = _load_ckpt(checkpoint_path, map_location)
        hp = dict(ckpt.get("hyper_parameters", {}))
        hp.pop("data_module_cls", None)
        hp.update({k: v for k, v in overrides.items() if k not in ("kwargs",)})
        model = cls(**hp)
        model.load_state_dict(ckpt["state_dict"], strict=True)
        ema = ckpt.get("ema")
        if ema is not None and ema.get("shadow_params") is not None:      # swapped in by eval() (model.py:517-531)
            model._ema_shadow = [(n, sh.detach().clone()) for n, sh in map_ema_shadow(model.named_parameters(), ema["shadow_params"])]
        return model

    def next_seed(self) -> int:
        self._calls += 1
        return (self._seed * 1000003 + self._calls) & 0xFFFFFFFFFFFF

    def forward_denoiser(self, y, **kwargs):
        return self.denoiser_net(y)                                   # model.py:663-665

    def forward_score(self, x, t, score_conditioning, sde_input=None, **kwargs):
        return -self.score_net(torch.cat([x] + list(score_conditioning), dim=1), t)    # model.py:654-661

    def get_pc_sampler(self, predictor_name, corrector_name, y, N=None, minibatch=None, scale_factor=None, conditioning=None, **kwargs):
        sde = self.sde.copy()
        sde.N = self.sde.N if N is None else N
        kwargs = {"eps": self.t_eps, **kwargs}
        return sampling.get_pc_sampler(predictor_name, corrector_name, sde=sde, score_fn=self.forward_score, y=y,
                                       conditioning=conditioning, **kwargs)

    def enhance_batch(self, ys, predictor="reverse_diffusion", corrector="none", N=30, corrector_steps=1, snr=0.5,
                      denoiser_only=False, seed: Optional[int] = None, utt_ids: Optional[Sequence[int]] = None,
                      return_nfe: bool = False, max_batch: int = 8):
        """Batched StoRM enhancement (the reference loops over files one at a time, enhancement.py:66-72).

        ``ys``: tensor [B, L] -> tensor [B, L]; or a list of utterances of any lengths -> list of [L_b] tensors (bucketed by
        padded frame count and run as ragged batches of at most ``max_batch``, like ``ScoreModel.enhance_batch``).  Per batch:
        per-utterance normalisation and STFT, ONE batched denoiser forward, ONE native N-step sampler around ``Y_denoised``
        with ``[Y, Y_denoised]`` as conditioning, batched iSTFT.  Noise is keyed per utterance id."""
        dev = next(self._nets()[0].parameters()).device
        if dev.type != "cuda":
            raise RuntimeError("StochasticRegenerationModel.enhance_batch needs the model on a CUDA device: call model.cuda() first")
        sp = self.data_module.spec_config()
        seed = self.next_seed() if seed is None else seed
        eng = (self.score_net or self.denoiser_net).engine()

        def run(yd, ids, lengths):
            norm = _native.wave_norm(yd, lengths)
            T_pad = _native.padded_frames(_native.num_frames(yd.shape[1], sp))
            Y = _native.stft(yd, norm, sp, T_pad, lengths=lengths).unsqueeze(1)
            Y_den = self.forward_denoiser(Y) if self.denoiser_net is not None else None
            nfe = 0
            if self.score_net is not None and not denoiser_only:
                cond = {"noisy": [Y], "post_denoiser": [Y_den], "both": [Y, Y_den]}[self.condition]
                sampler = self.get_pc_sampler(predictor, corrector, Y_den if Y_den is not None else Y, N=N,
                                              corrector_steps=corrector_steps, snr=snr, conditioning=cond, seed=seed, utt_ids=ids)
                sample, nfe = sampler()
            else:
                sample = Y_den
            return eng.istft(sample.squeeze(1), norm, yd.shape[1], sp, lengths=lengths), nfe

        if not torch.is_tensor(ys):
            ys = [y.reshape(-1) for y in ys]
            lens = [int(y.numel()) for y in ys]
            if len(set(lens)) > 1:
                from .parallel import bucket_utterances
                ids_all = list(range(len(ys))) if utt_ids is None else [int(u) for u in utt_ids]
                outs: List[Optional[torch.Tensor]] = [None] * len(ys)
                nfe = 0
                for idx in bucket_utterances(lens, sp.hop_length, max_batch):
                    Lmax = max(lens[i] for i in idx)
                    host = torch.zeros(len(idx), Lmax, dtype=torch.float32)
                    for r, i in enumerate(idx):
                        host[r, :lens[i]] = ys[i].to("cpu", torch.float32)
                    out, nfe = run(host.pin_memory().to(dev, non_blocking=True), [ids_all[i] for i in idx], [lens[i] for i in idx])
                    for r, i in enumerate(idx):
                        outs[i] = out[r, :lens[i]].to(ys[i].device)
                return (outs, nfe) if return_nfe else outs
            ys = torch.stack(ys)
        src = ys.device
        yd = ys.to(torch.float32)
        yd = yd.pin_memory().to(dev, non_blocking=True) if not yd.is_cuda else yd
        ids = list(range(yd.shape[0])) if utt_ids is None else [int(u) for u in utt_ids]
        out, nfe = run(yd, ids, None)
        out = out.to(src)
        return (out, nfe) if return_nfe else out

    def enhance(self, y, sampler_type="pc", predictor="reverse_diffusion", corrector="none", N=30, corrector_steps=1, snr=0.5,
                timeit=False, scale_factor=None, return_stft=False, denoiser_only=False, **kwargs):
        """model.py:720-780 for one utterance ``y`` [1, L]."""
        start = time.time()
        T_orig = y.size(1)
        dev = next(self._nets()[0].parameters()).device
        if dev.type != "cuda":
            raise RuntimeError("StochasticRegenerationModel.enhance needs the model on a CUDA device: call model.cuda() first")
        yd = y.to(dev, torch.float32)
        sp = self.data_module.spec_config()
        norm = _native.wave_norm(yd)
        T_pad = _native.padded_frames(_native.num_frames(T_orig, sp))
        Y = _native.stft(yd, norm, sp, T_pad).unsqueeze(1)
        nfe = 0
        Y_den = self.forward_denoiser(Y) if self.denoiser_net is not None else None
        if self.score_net is not None and not denoiser_only:
            cond = {"noisy": [Y], "post_denoiser": [Y_den], "both": [Y, Y_den]}[self.condition]
            if sampler_type != "pc":
                raise ValueError("{} is not a valid sampler type!".format(sampler_type))
            sampler = self.get_pc_sampler(predictor, corrector, Y_den, N=N, corrector_steps=corrector_steps, snr=snr,
                                          conditioning=cond, **kwargs)
            sample, nfe = sampler()
        else:
            sample = Y_den
        if return_stft:
            return sample.squeeze(), Y.squeeze(), T_orig, norm.item()
        eng = (self.score_net or self.denoiser_net).engine()
        x_hat = eng.istft(sample.squeeze(1), norm, T_orig, sp).squeeze().cpu()
        if timeit:
            return x_hat, nfe, (time.time() - start) / (len(x_hat) / 16000)
        return x_hat


def checkpoint_kind(path) -> str:
    """'storm' for StochasticRegenerationModel checkpoints (``denoiser_net.*`` / ``score_net.*`` keys), 'score' for
    ScoreModel / DiscriminativeModel ones (``dnn.*``)."""
    sd = _load_ckpt(path, "cpu").get("state_dict", {})
    if any(k.startswith(("denoiser_net.", "score_net.")) for k in sd):
        return "storm"
    return "score"


class _Tolerant:
    """Stand-in for classes pickled into reference checkpoints whose modules are absent here
    (the pickled ``data_module_cls`` hyper-parameter, model.py:61-62)."""

    def __init__(self, *a, **k):
        pass

    def __setstate__(self, state):
        pass


def _load_ckpt(path, map_location):
    import pickle

    class _U(pickle.Unpickler):
        def find_class(self, module, name):
            try:
                return super().find_class(module, name)
            except Exception:  # noqa: BLE001
                return _Tolerant

    class _P:
        Unpickler = _U
        load = staticmethod(pickle.load)
        __name__ = "pickle"

    return torch.load(path, map_location=map_location, pickle_module=_P, weights_only=False)
