"""storm_b200 - B200-native reverse-diffusion enhancement path of sp-uhh/storm.

Host code mirrors the reference's plugin surface (registries, ScoreModel.enhance, SpecsDataModule);
the work runs in hand-written sm_100a CUDA behind the C ABI of include/storm_b200.h."""
from .registry import Registry  # noqa: F401
from .backbones import BackboneRegistry  # noqa: F401
from .sdes import SDERegistry  # noqa: F401
from .sampling import PredictorRegistry, CorrectorRegistry  # noqa: F401
from .data_module import SpecsDataModule  # noqa: F401
from .model import ScoreModel, DiscriminativeModel, StochasticRegenerationModel  # noqa: F401

__version__ = "0.1.0"
