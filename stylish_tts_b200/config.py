"""Model configuration for the B200 engine.

Mirrors the fields of the reference's pydantic ``ModelConfig`` that the hot
path reads (reference: src/stylish_tts/lib/config_loader.py:375-433 and
src/stylish_tts/train/config/model.yml).  ``build_model`` only uses attribute
access, so either an object produced here or the reference's own pydantic
``ModelConfig`` instance can be handed to it.
"""
from __future__ import annotations

import io
import os
from types import SimpleNamespace
from typing import Any, Union

import yaml

_REQUIRED = {
    "": ["sample_rate", "n_mels", "n_fft", "win_length", "hop_length",
         "coarse_multiplier", "style_dim", "inter_dim"],
    "decoder": ["hidden_dim", "residual_dim"],
    "generator": ["input_dim", "io_conv_kernel_size", "conformer_layers", "conv_layers"],
    "text_encoder": ["tokens", "hidden_dim", "filter_channels", "heads", "layers",
                     "kernel_size", "dropout"],
    "style_encoder": ["n_mels", "n_fft", "win_length", "hop_length", "max_channels",
                      "skip_downsample"],
    "duration_predictor": ["n_layer", "duration_classes", "max_duration", "dropout",
                           "last_dropout"],
    "pitch_energy_predictor": ["inter_dim", "dropout"],
}


class ConfigError(ValueError):
    pass


def _ns(obj: Any) -> Any:
    if isinstance(obj, dict):
        return SimpleNamespace(**{k: _ns(v) for k, v in obj.items()})
    return obj


def _validate(raw: dict) -> None:
    for section, keys in _REQUIRED.items():
        node = raw if section == "" else raw.get(section)
        if not isinstance(node, dict):
            raise ConfigError(f"model config: missing section '{section}'")
        for k in keys:
            if k not in node:
                where = section or "<root>"
                raise ConfigError(f"model config: missing field '{k}' in {where}")


def load_model_config_yaml(src: Union[str, io.IOBase]) -> SimpleNamespace:
    """Load a ``model.yml``.  Accepts a path or an open stream, like the
    reference's loader (config_loader.py:467-480)."""
    if isinstance(src, (str, os.PathLike)):
        with open(src, "r", encoding="utf-8") as f:
            raw = yaml.safe_load(f)
    else:
        raw = yaml.safe_load(src)
    _validate(raw)
    return _ns(raw)


def default_model_config() -> SimpleNamespace:
    here = os.path.dirname(os.path.abspath(__file__))
    return load_model_config_yaml(os.path.join(here, "model.yml"))
