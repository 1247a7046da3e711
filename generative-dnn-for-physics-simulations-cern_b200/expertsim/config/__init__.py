"""Configuration loading.  Replaces the Hydra/OmegaConf ``load_config`` of the reference (cli.py:58-97) with a
PyYAML-based loader that yields the same attribute-style, mutable, nested config (``cfg.model.router.tau_start``,
``**cfg.model.generator``) and accepts the same ``key=value`` dotted overrides.  The YAML schema is the reference's
(expertsim/config/default.yaml:1-58)."""
from __future__ import annotations

import os
import re
from typing import Iterable, Optional

import yaml

DEFAULT_CONFIG = os.path.join(os.path.dirname(os.path.abspath(__file__)), "default.yaml")
_FLOAT = re.compile(r"^[+-]?(\d+\.?\d*|\.\d+)([eE][+-]?\d+)?$")


def _scalar(v):
    """OmegaConf-style scalar typing: PyYAML leaves '1e-4' a string, OmegaConf makes it a float."""
    if isinstance(v, str):
        s = v.strip()
        if s.lower() in ("null", "none", "~"):
            return None
        if s.lower() in ("true", "false"):
            return s.lower() == "true"
        if re.fullmatch(r"[+-]?\d+", s):
            return int(s)
        if _FLOAT.match(s):
            return float(s)
    return v


class Config(dict):
    """dict with attribute access, recursive; new keys may be added freely (the reference runs with struct off)."""

    def __init__(self, data=None):
        super().__init__()
        for k, v in (data or {}).items():
            self[k] = v

    @staticmethod
    def _wrap(v):
        if isinstance(v, dict) and not isinstance(v, Config):
            return Config(v)
        if isinstance(v, (list, tuple)):
            return [Config._wrap(x) for x in v]
        return _scalar(v)

    def __setitem__(self, k, v):
        super().__setitem__(k, Config._wrap(v))

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError:
            raise AttributeError(k) from None

    def __setattr__(self, k, v):
        self[k] = v

    def set_path(self, dotted: str, value):
        *path, leaf = dotted.split(".")
        node = self
        for p in path:
            if p not in node or not isinstance(node[p], Config):
                node[p] = Config()
            node = node[p]
        node[leaf] = value

    def to_dict(self):
        return {k: (v.to_dict() if isinstance(v, Config) else v) for k, v in self.items()}


def load_config(config_path: Optional[str] = None, overrides: Optional[Iterable[str]] = None) -> Config:
    """YAML file (default: the packaged default.yaml) + ``a.b.c=value`` overrides -> Config."""
    path = config_path or DEFAULT_CONFIG
    if not os.path.exists(path) and os.path.exists(path + ".yaml"):
        path += ".yaml"
    with open(path) as f:
        cfg = Config(yaml.safe_load(f) or {})
    for ov in overrides or ():
        if "=" not in ov:
            raise ValueError(f"override '{ov}' is not of the form key=value")
        k, v = ov.split("=", 1)
        cfg.set_path(k.lstrip("+"), yaml.safe_load(v))
    return cfg
