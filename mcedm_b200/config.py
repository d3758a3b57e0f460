"""Hydra-free loader for the reference's YAML config tree.

The reference is driven by Hydra (`run.py:30`, `configs/config_adm_edm_mcedm_res32.yaml`); neither
hydra nor omegaconf exist in this image, so the handful of features the hot path needs are
re-implemented on PyYAML: `defaults:` list composition, `_target_` passthrough, dotted CLI overrides
(`datamodule.batch_size=16 diff_sampler.n_samples=1`, README.md:19) and an attribute dictionary with
OmegaConf's semantics (missing key -> AttributeError, so the reference's `hasattr(hparams.model, ..)`
feature probes work; `.get(key, default)`).  The reference's own `utils.DotDict` is NOT suitable: it
raises KeyError on missing attributes (utils.py:4-8), which breaks `hasattr`.
"""
from __future__ import annotations

import copy
import os
from typing import Any, Iterable

import yaml

_PKG_CONFIGS = os.path.join(os.path.dirname(os.path.abspath(__file__)), "configs")


class AttrDict(dict):
    """dict with attribute access; missing attribute raises AttributeError (like OmegaConf's DictConfig)."""

    def __getattr__(self, name: str) -> Any:
        try:
            return self[name]
        except KeyError:
            raise AttributeError(name) from None

    def __setattr__(self, name: str, value: Any) -> None:
        self[name] = value

    def __delattr__(self, name: str) -> None:
        try:
            del self[name]
        except KeyError:
            raise AttributeError(name) from None

    def __deepcopy__(self, memo):
        return AttrDict({k: copy.deepcopy(v, memo) for k, v in self.items()})


def to_attr(obj: Any) -> Any:
    if isinstance(obj, dict):
        return AttrDict({k: to_attr(v) for k, v in obj.items()})
    if isinstance(obj, (list, tuple)):
        return [to_attr(v) for v in obj]
    return obj


def load_yaml(path: str) -> AttrDict:
    with open(path, "r") as f:
        return to_attr(yaml.safe_load(f) or {})


def _set_dotted(cfg: dict, dotted: str, value: Any) -> None:
    keys = dotted.split(".")
    node = cfg
    for k in keys[:-1]:
        if k not in node or not isinstance(node[k], dict):
            node[k] = AttrDict()
        node = node[k]
    node[keys[-1]] = value


def compose(config_name: str, overrides: Iterable[str] = (), config_dir: str | None = None) -> AttrDict:
    """Compose `<config_dir>/<config_name>.yaml` the way Hydra would for this repo's configs.

    Only what the m-cedm configs use is supported: a `defaults:` list of `{group: option}` entries
    (entries starting with `override hydra/` and `_self_` are skipped), group files under
    `<config_dir>/<group>/<option>.yaml` placed at key `<group>`, and `key.sub=value` overrides parsed
    with YAML scalar rules.  `group=option` overrides re-select a defaults entry.
    """
    config_dir = config_dir or _PKG_CONFIGS
    name = config_name[:-5] if config_name.endswith(".yaml") else config_name
    root = load_yaml(os.path.join(config_dir, name + ".yaml"))
    defaults = root.pop("defaults", [])
    overrides = list(overrides)
    group_choice = {}
    for ov in overrides:
        k, _, v = ov.partition("=")
        if "." not in k and os.path.isdir(os.path.join(config_dir, k)):
            group_choice[k] = v
    cfg = AttrDict()
    for entry in defaults:
        if entry == "_self_" or not isinstance(entry, dict):
            continue
        for group, option in entry.items():
            if group.startswith("override "):
                continue
            option = group_choice.get(group, option)
            cfg[group] = load_yaml(os.path.join(config_dir, group, f"{option}.yaml"))
    for k, v in root.items():
        if k == "hydra":
            continue
        cfg[k] = v
    for ov in overrides:
        k, _, v = ov.partition("=")
        if k in group_choice and "." not in k:
            continue
        _set_dotted(cfg, k, to_attr(yaml.safe_load(v)))
    return cfg


def default_config_dir() -> str:
    return _PKG_CONFIGS
