"""Import the UNMODIFIED reference (sp-uhh/storm at /root/reference) on CPU under stub modules.

TEST INFRASTRUCTURE ONLY.  This file is used by ``oracle/gen_golden.py`` in the build container to
produce the committed fixtures under ``tests/golden/``; it is never imported by the product path and it
cannot run on the GPU box (``/root/reference`` does not exist there).

The reference's hot path imports fine once four third-party modules that are not
installed here are stubbed: ``pydub``, ``matplotlib(.pyplot)``, ``h5py``, ``pytorch_lightning``.
Importing ``sgmse.backbones`` JIT-builds the reference's two CUDA extensions
(``sgmse/backbones/ncsnpp_utils/op/upfirdn2d.py:12-18``); on CPU tensors they are never called
(``op/upfirdn2d.py:147-151`` dispatches to ``upfirdn2d_native``).
"""
import os
import sys
import types

REF_ROOT = os.environ.get("STORM_REFERENCE_ROOT", "/root/reference")


def available() -> bool:
    return os.path.isdir(os.path.join(REF_ROOT, "sgmse"))


def _stub(name, **attrs):
    m = types.ModuleType(name)
    for k, v in attrs.items():
        setattr(m, k, v)
    sys.modules[name] = m
    return m


def load():
    """Return a namespace with the reference classes used on the hot path."""
    if not available():
        raise RuntimeError(f"reference tree not found at {REF_ROOT}")
    import torch

    os.environ.setdefault("TORCH_EXTENSIONS_DIR", "/tmp/storm_ref_ext")
    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)
    if "pydub" not in sys.modules:
        _stub("pydub", AudioSegment=object)
    if "matplotlib" not in sys.modules:
        plt = _stub("matplotlib.pyplot")
        _stub("matplotlib", pyplot=plt)
    if "h5py" not in sys.modules:
        _stub("h5py")
    if "pytorch_lightning" not in sys.modules:
        _stub("pytorch_lightning", LightningDataModule=object, LightningModule=torch.nn.Module)

    from sgmse.backbones import BackboneRegistry
    from sgmse.backbones.ncsnpp_utils import layerspp, up_or_down_sampling
    from sgmse.sdes import SDERegistry, OUVESDE
    from sgmse import sampling
    from sgmse.sampling import predictors, correctors
    from sgmse.data_module import SpecsDataModule
    from sgmse.util.other import pad_spec, si_sdr, si_sdr_torch

    ns = types.SimpleNamespace(
        BackboneRegistry=BackboneRegistry, layerspp=layerspp, up_or_down_sampling=up_or_down_sampling,
        SDERegistry=SDERegistry, OUVESDE=OUVESDE, sampling=sampling, predictors=predictors,
        correctors=correctors, SpecsDataModule=SpecsDataModule, pad_spec=pad_spec, si_sdr=si_sdr, si_sdr_torch=si_sdr_torch,
    )
    return ns
