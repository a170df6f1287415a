"""Config loading with the reference's Hydra/OmegaConf conventions (configs/exp/gpv.yaml): YAML, `${a.b}`
interpolation, dotted `key=value` overrides, attribute access and `.items()` (losses.py:147 iterates `cfg.items()`).
Uses hydra/omegaconf objects unchanged when the caller passes them; this loader exists because neither is a
dependency of the hot path."""
import os
import re

import yaml

DEFAULT_YAML = os.path.join(os.path.dirname(os.path.abspath(__file__)), "configs", "gpv.yaml")


class Cfg(dict):
    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:
            raise AttributeError(k) from e

    def __setattr__(self, k, v):
        self[k] = v


def _wrap(o):
    if isinstance(o, dict):
        return Cfg({k: _wrap(v) for k, v in o.items()})
    if isinstance(o, list):
        return [_wrap(v) for v in o]
    return o


_NUM = re.compile(r"^[+-]?(\d+\.?\d*|\.\d+)([eE][+-]?\d+)?$")


def _coerce(v):
    if isinstance(v, str) and _NUM.match(v):
        if re.fullmatch(r"[+-]?\d+", v):
            return int(v)
        return float(v)
    return v


def _get(root, dotted):
    cur = root
    for p in dotted.split("."):
        cur = cur[p]
    return cur


def _resolve(node, root):
    if isinstance(node, dict):
        return {k: _resolve(v, root) for k, v in node.items()}
    if isinstance(node, list):
        return [_resolve(v, root) for v in node]
    if isinstance(node, str):
        m = re.fullmatch(r"\$\{([^}]+)\}", node)
        if m:
            return _resolve(_get(root, m.group(1)), root)
        return _coerce(re.sub(r"\$\{([^}]+)\}", lambda mm: str(_resolve(_get(root, mm.group(1)), root)), node))
    return node


def load_config(path=None, overrides=None):
    """overrides: dict or list of 'a.b=value' strings (Hydra command-line style)."""
    with open(path or DEFAULT_YAML) as f:
        raw = yaml.safe_load(f)
    raw.pop("defaults", None)
    raw.pop("hydra", None)
    if isinstance(overrides, (list, tuple)):
        overrides = dict(o.split("=", 1) for o in overrides)
    for k, v in (overrides or {}).items():
        cur = raw
        parts = k.split(".")
        for p in parts[:-1]:
            cur = cur.setdefault(p, {})
        cur[parts[-1]] = yaml.safe_load(v) if isinstance(v, str) and not v.startswith("/") else v
    return _wrap(_resolve(raw, raw))
