"""Loader for the UNMODIFIED reference (Stylish-TTS) as a Python import.

TEST INFRASTRUCTURE ONLY.  Works only where /root/reference is mounted (the
build container); the GPU box never imports this.  Used by
tests/golden/make_golden.py to generate committed fixtures and by the
`-m "not gpu"` tests to pin oracle/ against the live reference when present.

Follows SURVEY.md Appendix A: bypass stylish_tts/__init__.py (it imports the
ONNX CLI), stub munch / matplotlib / accelerate / soundfile / librosa.
"""
import os
import sys
import types

REF_ROOT = os.environ.get("STYLISH_REF_ROOT", "/root/reference")
REF_SRC = os.path.join(REF_ROOT, "src")


def available() -> bool:
    return os.path.isdir(os.path.join(REF_SRC, "stylish_tts", "train", "models"))


class Munch(dict):
    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError:
            raise AttributeError(k)

    def __setattr__(self, k, v):
        self[k] = v


def _stub(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


_loaded = False


def load():
    """Make `import stylish_tts.train...` resolve to the reference sources."""
    global _loaded
    if _loaded:
        return
    if not available():
        raise RuntimeError(f"reference sources not found under {REF_SRC}")
    import torch  # noqa: F401
    import torchaudio  # noqa: F401
    import transformers  # noqa: F401  (must precede the accelerate stub)

    if REF_SRC not in sys.path:
        sys.path.insert(0, REF_SRC)
    if "munch" not in sys.modules:
        _stub("munch", Munch=Munch)
    if "matplotlib" not in sys.modules:
        plt = _stub("matplotlib.pyplot", Figure=object)
        gs = _stub("matplotlib.gridspec")
        _stub("matplotlib", pyplot=plt, gridspec=gs)

    class _Any:
        def __init__(self, *a, **k):
            pass

    for name in ("accelerate", "accelerate.accelerator"):
        if name not in sys.modules:
            _stub(name, Accelerator=_Any, DistributedDataParallelKwargs=_Any)
    if "soundfile" not in sys.modules:
        _stub("soundfile")
    if "librosa" not in sys.modules:
        lf = _stub("librosa.filters", mel=None)
        _stub("librosa", filters=lf)
    pkg = types.ModuleType("stylish_tts")
    pkg.__path__ = [os.path.join(REF_SRC, "stylish_tts")]
    pkg.__file__ = os.path.join(REF_SRC, "stylish_tts", "__init__.py")
    sys.modules["stylish_tts"] = pkg
    _loaded = True


def model_config():
    load()
    from stylish_tts.lib.config_loader import load_model_config_yaml

    path = os.path.join(REF_SRC, "stylish_tts", "train", "config", "model.yml")
    with open(path) as f:
        return load_model_config_yaml(f)


def build_model():
    load()
    from stylish_tts.train.models.models import build_model as _bm

    return _bm(model_config())
