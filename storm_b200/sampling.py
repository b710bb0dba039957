"""Sampler plugin boundary: ``PredictorRegistry`` / ``CorrectorRegistry`` and the sampler factories.

Surface of sgmse/sampling/__init__.py:27-141, predictors.py:9-80, correctors.py:8-106:
same registry names (``reverse_diffusion``, ``euler_maruyama``, ``none`` / ``ald``, ``langevin``, ``none``),
same factory signatures and return convention ``(sample, nfe)``.

When ``score_fn`` is a :class:`storm_b200.model.ScoreModel` (the only caller in the reference,
model.py:202-244) the whole N-step loop runs inside the native engine (``storm_pc_sample``): score network,
SDE algebra and noise injection are CUDA kernels, no per-step Python.  For any other ``score_fn`` the
classes below run the same update rules as plain tensor expressions so third-party code written against
the reference API keeps working.
"""
from __future__ import annotations

import abc

import numpy as np
import torch

from . import _native, sdes
from .registry import Registry

PredictorRegistry = Registry("Predictor")
CorrectorRegistry = Registry("Corrector")

__all__ = ["PredictorRegistry", "CorrectorRegistry", "Predictor", "Corrector", "get_pc_sampler", "get_ode_sampler"]


class Predictor(abc.ABC):
    def __init__(self, sde, score_fn, probability_flow=False):
        self.sde, self.score_fn, self.probability_flow = sde, score_fn, probability_flow
        self.rsde = sde.reverse(score_fn)      # the reference drops probability_flow here too (predictors.py:15-20)

    @abc.abstractmethod
    def update_fn(self, x, t, *args):
        ...


@PredictorRegistry.register("euler_maruyama")
class EulerMaruyamaPredictor(Predictor):
    def update_fn(self, x, t, *args, **kwargs):
        dt = -1.0 / self.rsde.N
        z = torch.randn_like(x)
        f, g = self.rsde.sde(x, t, *args)
        x_mean = x + f * dt
        return x_mean + sdes._bcast(g, x) * float(np.sqrt(-dt)) * z, x_mean


@PredictorRegistry.register("reverse_diffusion")
class ReverseDiffusionPredictor(Predictor):
    def update_fn(self, x, t, *args, **kwargs):
        f, g = self.rsde.discretize(x, t, *args)
        z = torch.randn_like(x)
        x_mean = x - f
        return x_mean + sdes._bcast(g, x) * z, x_mean


@PredictorRegistry.register("none")
class NonePredictor(Predictor):
    def __init__(self, *args, **kwargs):
        pass

    def update_fn(self, x, t, *args, **kwargs):
        return x, x


class Corrector(abc.ABC):
    def __init__(self, sde, score_fn, snr, n_steps):
        self.sde, self.score_fn, self.snr, self.n_steps = sde, score_fn, snr, n_steps

    @abc.abstractmethod
    def update_fn(self, x, t, *args):
        ...


@CorrectorRegistry.register("langevin")
class LangevinCorrector(Corrector):
    def update_fn(self, x, t, *args, **kwargs):
        x_mean = x
        for _ in range(self.n_steps):
            grad = self.score_fn(x, t, *args)
            noise = torch.randn_like(x)
            gn = torch.norm(grad.reshape(grad.shape[0], -1), dim=-1).mean()
            zn = torch.norm(noise.reshape(noise.shape[0], -1), dim=-1).mean()
            step = (self.snr * zn / gn) ** 2 * 2
            x_mean = x + step * grad
            x = x_mean + noise * torch.sqrt(step * 2)
        return x, x_mean


@CorrectorRegistry.register("ald")
class AnnealedLangevinDynamics(Corrector):
    def __init__(self, sde, score_fn, snr, n_steps):
        super().__init__(sde, score_fn, snr, n_steps)
        if not isinstance(sde, sdes.OUVESDE):
            raise NotImplementedError(f"SDE class {sde.__class__.__name__} not yet supported.")

    def update_fn(self, x, t, *args, **kwargs):
        std = self.sde.marginal_prob(x, t, *args)[1]
        x_mean = x
        for _ in range(self.n_steps):
            grad = self.score_fn(x, t, *args)
            noise = torch.randn_like(x)
            step = sdes._bcast((self.snr * std) ** 2 * 2, x)
            x_mean = x + step * grad
            x = x_mean + noise * torch.sqrt(step * 2)
        return x, x_mean


@CorrectorRegistry.register("none")
class NoneCorrector(Corrector):
    def __init__(self, *args, **kwargs):
        self.snr, self.n_steps = 0, 0

    def update_fn(self, x, t, *args, **kwargs):
        return x, x


def _native_model(score_fn):
    """(model, backbone) when ``score_fn`` is one of this package's models (or its stock score method): only then may the
    native engine replace the Python loop.  A subclass that overrides ``forward`` / ``forward_score`` (guidance, wrappers)
    or any other callable keeps the generic Python sampler below, which calls it like the reference does."""
    from .model import DiscriminativeModel, ScoreModel, StochasticRegenerationModel   # local import: model imports this module
    owner = getattr(score_fn, "__self__", score_fn)
    func = getattr(score_fn, "__func__", None)
    if isinstance(owner, ScoreModel) and not isinstance(owner, DiscriminativeModel):
        if type(owner).forward is ScoreModel.forward and func in (None, ScoreModel.forward):
            return owner, owner.dnn
        return None
    if isinstance(owner, StochasticRegenerationModel):
        if type(owner).forward_score is StochasticRegenerationModel.forward_score and func in (None, StochasticRegenerationModel.forward_score):
            return owner, owner.score_net
    return None


def get_pc_sampler(predictor_name, corrector_name, sde, score_fn, y, denoise=True, eps=3e-2, snr=0.1, corrector_steps=1,
                   probability_flow: bool = False, conditioning=None, intermediate=False, seed=None, utt_ids=None,
                   noise=None, **kwargs):
    """Predictor-corrector sampler factory (sampling/__init__.py:27-68).  Extensions: ``seed`` / ``utt_ids``
    key the counter-based noise (sharding-invariant), ``noise`` injects pre-drawn complex normals for parity
    tests, ``probability_flow=True`` really selects the fixed-step PF update (a no-op in the reference)."""
    predictor_cls = PredictorRegistry.get_by_name(predictor_name)
    corrector_cls = CorrectorRegistry.get_by_name(corrector_name)
    nat = _native_model(score_fn)
    if nat is not None:
        model, dnn = nat
        if not isinstance(sde, sdes.OUVESDE):
            raise NotImplementedError(f"SDE class {sde.__class__.__name__} not yet supported.")
        sc = _native.make_sampler_config(predictor_name, corrector_name, sde.N, corrector_steps, snr, eps, sde.theta,
                                         sde.sigma_min, sde.sigma_max, denoise, probability_flow)

        def native_pc_sampler():
            s = model.next_seed() if seed is None else seed
            # conditioning = list of spectrograms stacked after x (StoRM, model.py:742-756); default [y]
            return dnn.engine().pc_sample(y, sc, seed=s, utt_ids=utt_ids, noise=noise, cond=conditioning)

        return native_pc_sampler

    # generic path: a conditional score function receives `conditioning=` exactly as in the reference
    # (sampling/__init__.py:60-63: update_fn(xt, vec_t, y, conditioning=conditioning) -> score_fn(x, t, y, conditioning=...))
    fn = score_fn if conditioning is None else (lambda x, t, *a, **k: score_fn(x, t, *a, conditioning=conditioning, **k))
    predictor = predictor_cls(sde, fn, probability_flow=probability_flow)
    corrector = corrector_cls(sde, fn, snr=snr, n_steps=corrector_steps)

    def pc_sampler():
        with torch.no_grad():
            xt = sde.prior_sampling(y.shape, y).to(y.device)
            xt_mean = xt
            for t in torch.linspace(sde.T, eps, sde.N, device=y.device):
                vec_t = torch.ones(y.shape[0], device=y.device) * t
                xt, xt_mean = corrector.update_fn(xt, vec_t, y)
                xt, xt_mean = predictor.update_fn(xt, vec_t, y)
            return (xt_mean if (denoise and sde.N) else xt), sde.N * (corrector.n_steps + 1)

    return pc_sampler


def get_ode_sampler(sde, score_fn, y, inverse_scaler=None, denoise=True, rtol=1e-5, atol=1e-5, method="RK45", eps=3e-2,
                    device="cuda", seed=None, utt_ids=None, noise=None, **kwargs):
    """Probability-flow ODE sampler with scipy's adaptive solver on the host, like the reference
    (sampling/__init__.py:71-141): every right-hand-side evaluation is one native drift call
    (score network + drift algebra on the GPU); the state crosses PCIe per evaluation exactly as it does there."""
    from scipy import integrate

    nat = _native_model(score_fn)
    model = nat[0] if nat is not None and hasattr(nat[0], "dnn") else None      # ScoreModel only (as model.py:224-244)
    rsde = sde.reverse(score_fn, probability_flow=True)
    sc = None
    if model is not None:
        sc = _native.make_sampler_config("reverse_diffusion", "none", sde.N, 0, 0.0, eps, sde.theta, sde.sigma_min, sde.sigma_max)

    def drift(x, t):
        if model is not None:
            return model.dnn.engine().ode_drift(x, y, float(t), sc)
        return rsde.sde(x, torch.ones(y.shape[0], device=y.device) * t, y)[0]

    def ode_sampler(z=None, **kw):
        with torch.no_grad():
            if noise is not None:
                z0, rest = noise[0], noise[1:]
            else:
                z0, rest = None, None
            std = sde._std(torch.ones((y.shape[0],), device=y.device))[:, None, None, None]
            if z0 is None:
                if model is not None:
                    s = model.next_seed() if seed is None else seed
                    ids = list(range(y.shape[0])) if utt_ids is None else utt_ids
                    z0 = torch.stack([_native.philox_normal(s, u, 0, y[0].numel()).reshape(y.shape[1:]) for u in ids])
                else:
                    z0 = torch.randn_like(y)
            x = y + z0 * std

            def rhs(t, xf):
                xx = torch.from_numpy(xf.reshape(tuple(y.shape))).to(y.device).type(torch.complex64)
                return drift(xx, t).detach().cpu().numpy().reshape((-1,))

            sol = integrate.solve_ivp(rhs, (sde.T, eps), x.detach().cpu().numpy().reshape((-1,)), rtol=rtol, atol=atol, method=method)
            x = torch.tensor(sol.y[:, -1]).reshape(y.shape).to(y.device).type(torch.complex64)
            if denoise:   # one noise-free reverse-diffusion step at t = eps with dt = 1/N (sampling/__init__.py:97-100)
                vec_eps = torch.ones(y.shape[0], device=y.device) * eps
                f, _ = sde.reverse(score_fn).discretize(x, vec_eps, y)
                x = x - f
            if inverse_scaler is not None:
                x = inverse_scaler(x)
            return x, sol.nfev

    return ode_sampler
