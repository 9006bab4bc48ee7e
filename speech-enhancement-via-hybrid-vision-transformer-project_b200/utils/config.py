"""YAML configuration loading with the reference's merge order and semantics (reference utils/config.py:13-110):
data_config.yaml, then model_config.yaml, then train_config.yaml, later files overriding earlier ones key by key."""
from __future__ import annotations

from pathlib import Path
from typing import Any, Dict, Union

import yaml


def load_config(config_path: Union[str, Path]) -> Dict[str, Any]:
    path = Path(config_path)
    if not path.exists():
        raise FileNotFoundError(f"Configuration file not found: {path}")
    with open(path, "r") as f:
        return yaml.safe_load(f) or {}


def merge_configs(base: Dict[str, Any], override: Dict[str, Any]) -> Dict[str, Any]:
    out = dict(base)
    for k, v in override.items():
        out[k] = merge_configs(out[k], v) if isinstance(v, dict) and isinstance(out.get(k), dict) else v
    return out


def load_all_configs(config_dir: str = "config") -> Dict[str, Any]:
    d = Path(config_dir)
    if not d.exists():
        raise FileNotFoundError(f"Configuration directory not found: {d}")
    merged: Dict[str, Any] = {}
    for name in ("data_config.yaml", "model_config.yaml", "train_config.yaml"):
        if (d / name).exists():
            merged = merge_configs(merged, load_config(d / name))
        else:
            print(f"Warning: Configuration file not found: {d / name}")
    return merged
