"""Imports the UNMODIFIED reference (katehai/m-cedm) from /root/reference in the build container.

Used only by make_golden.py (fixture generation) and by CPU tests that cross-check the oracle against
the live reference when the mount is present.  Nothing here is read on the GPU box, where
/root/reference does not exist.

The reference needs two packages this image lacks; both are replaced by inert stand-ins that carry no
numerics:  `pytorch_lightning` (LightningModule -> nn.Module with no-op logging) and `h5py` (never
called: the datasets are instantiated with object.__new__ to reach their mask logic only).
"""
from __future__ import annotations

import os
import sys
import types

import torch

REF_ROOT = os.environ.get("MCEDM_REFERENCE_ROOT", "/root/reference")


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REF_ROOT, "models", "mcedm.py"))


def _install_stubs():
    if "pytorch_lightning" not in sys.modules:
        pl = types.ModuleType("pytorch_lightning")

        class LightningModule(torch.nn.Module):
            def __init__(self, *a, **k):
                super().__init__()
                self.logged = {}
                self.current_epoch = 0
                self.trainer = None

            def save_hyperparameters(self, *a, **k):
                pass

            def log(self, name, value, **kw):
                self.logged[name] = value

            def optimizer_step(self, *a, **k):
                pass

        class LightningDataModule:
            def __init__(self, *a, **k):
                pass

        class Callback:
            pass

        pl.LightningModule = LightningModule
        pl.LightningDataModule = LightningDataModule
        pl.Callback = Callback
        pl.seed_everything = lambda seed, workers=False: torch.manual_seed(seed)
        sys.modules["pytorch_lightning"] = pl
    if "h5py" not in sys.modules:
        sys.modules["h5py"] = types.ModuleType("h5py")


def import_reference():
    """Returns a namespace with the reference modules needed for the hot path."""
    if not reference_available():
        raise RuntimeError(f"reference not found under {REF_ROOT}")
    _install_stubs()
    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)
    import importlib

    ns = types.SimpleNamespace()
    ns.adm_blocks = importlib.import_module("models.adm_blocks")
    ns.mcedm = importlib.import_module("models.mcedm")
    ns.ddim = importlib.import_module("models.ddim")
    ns.losses = importlib.import_module("models.losses")
    ns.h5_dataset = importlib.import_module("datamodules.h5_dataset")
    return ns


def reference_hparams(config_name: str):
    """hparams tree of a reference experiment, composed from the reference's own YAML files."""
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
    from mcedm_b200.config import compose

    cfg = compose(config_name, config_dir=os.path.join(REF_ROOT, "configs"))
    return cfg
