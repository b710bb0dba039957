"""SDE plugin boundary: ``SDERegistry`` with the Ornstein-Uhlenbeck variance-exploding SDE.

Surface of sgmse/sdes.py:17-252 (``ouve``).  ``ouvp`` (sdes.py:255-327) is outside the path
(SURVEY.md section 2, row 8).  The tensor methods below are thin torch expressions kept for API
compatibility (custom samplers written against the reference keep working on any device); the shipped
samplers evaluate the same formulas inside the native engine (csrc/engine.cu: ouve_std / ouve_diffusion).
"""
from __future__ import annotations

import abc
import warnings

import numpy as np
import torch

from .registry import Registry

SDERegistry = Registry("SDE")


class SDE(abc.ABC):
    def __init__(self, N):
        self.N = N

    @property
    @abc.abstractmethod
    def T(self):
        ...

    @abc.abstractmethod
    def sde(self, x, t, *args):
        ...

    @abc.abstractmethod
    def marginal_prob(self, x, t, *args):
        ...

    @abc.abstractmethod
    def prior_sampling(self, shape, *args):
        ...

    @abc.abstractmethod
    def copy(self):
        ...

    def discretize(self, x, t, *args):
        """Euler-Maruyama discretisation f = drift/N, G = diffusion*sqrt(1/N) (sdes.py:73-90)."""
        dt = 1 / self.N
        drift, diffusion = self.sde(x, t, *args)
        return drift * dt, diffusion * float(np.sqrt(dt))

    def reverse(self, score_model, probability_flow: bool = False):
        """Reverse-time SDE/ODE built around ``score_model`` (sdes.py:92-159)."""
        return _Reverse(self, score_model, probability_flow)


def _bcast(v, x):
    return v.view(*v.shape, *((1,) * (x.ndim - v.ndim))) if torch.is_tensor(v) and v.ndim < x.ndim else v


class _Reverse:
    def __init__(self, fwd: SDE, score_model, probability_flow: bool):
        self.fwd, self.score_model, self.probability_flow = fwd, score_model, probability_flow
        self.N = fwd.N

    @property
    def T(self):
        return self.fwd.T

    def rsde_parts(self, x, t, *args, **kwargs):
        drift, diffusion = self.fwd.sde(x, t, *args)
        score = self.score_model(x, t, *args)
        diffusion = _bcast(diffusion, x)
        score_drift = -diffusion ** 2 * score * (0.5 if self.probability_flow else 1.0)
        return dict(total_drift=drift + score_drift, diffusion=torch.zeros_like(diffusion) if self.probability_flow else diffusion,
                    sde_drift=drift, sde_diffusion=diffusion, score_drift=score_drift, score=score)

    def sde(self, x, t, *args, **kwargs):
        p = self.rsde_parts(x, t, *args, **kwargs)
        return p["total_drift"], p["diffusion"]

    def discretize(self, x, t, *args, **kwargs):
        f, G = self.fwd.discretize(x, t, *args)
        G = _bcast(G, x)
        rev_f = f - G ** 2 * self.score_model(x, t, *args) * (0.5 if self.probability_flow else 1.0)
        return rev_f, (torch.zeros_like(G) if self.probability_flow else G)


@SDERegistry.register("ouve")
class OUVESDE(SDE):
    """dx = theta (y - x) dt + sigma_min (sigma_max/sigma_min)^t sqrt(2 log(sigma_max/sigma_min)) dw (sdes.py:166-252)."""

    def __init__(self, theta=1.5, sigma_min=0.05, sigma_max=0.5, N=1000, **ignored_kwargs):
        super().__init__(N)
        self.theta, self.sigma_min, self.sigma_max = theta, sigma_min, sigma_max
        self.logsig = float(np.log(sigma_max / sigma_min))

    @staticmethod
    def add_argparse_args(parser):
        parser.add_argument("--sde-n", type=int, default=1000, help="The number of timesteps in the SDE discretization. 1000 by default")
        parser.add_argument("--theta", type=float, default=1.5, help="The constant stiffness of the Ornstein-Uhlenbeck process.")
        parser.add_argument("--sigma-min", type=float, default=0.05, help="The minimum sigma to use.")
        parser.add_argument("--sigma-max", type=float, default=0.5, help="The maximum sigma to use.")
        return parser

    def copy(self):
        return OUVESDE(self.theta, self.sigma_min, self.sigma_max, N=self.N)

    @property
    def T(self):
        return 1

    def sde(self, x, t, y):
        sigma = self.sigma_min * (self.sigma_max / self.sigma_min) ** t
        return self.theta * (y - x), sigma * float(np.sqrt(2 * self.logsig))

    def _mean(self, x0, t, y):
        w = torch.exp(-self.theta * t)[:, None, None, None]
        return w * x0 + (1 - w) * y

    def _std(self, t, **kwargs):
        th, ls = self.theta, self.logsig
        return torch.sqrt(self.sigma_min ** 2 * torch.exp(-2 * th * t) * (torch.exp(2 * (th + ls) * t) - 1) * ls / (th + ls))

    def marginal_prob(self, x0, t, y):
        return self._mean(x0, t, y), self._std(t)

    def prior_sampling(self, shape, y):
        if shape != y.shape:
            warnings.warn(f"Target shape {shape} does not match shape of y {y.shape}! Ignoring target shape.")
        std = self._std(torch.ones((y.shape[0],), device=y.device))
        return y + torch.randn_like(y) * std[:, None, None, None]

    def prior_logp(self, z):
        raise NotImplementedError("prior_logp for OU SDE not yet implemented!")
