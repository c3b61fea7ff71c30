"""stylish_tts_b200 — B200-native (sm_100a) engine for the Stylish-TTS
forward / training hot path.  See DESIGN.md for scope and INTEGRATION.md for
how it plugs into the reference's train.py / inference graph."""
from .config import default_model_config, load_model_config_yaml  # noqa: F401
from .modules import (DecoderPrediction, DurationPredictor, DurationProcessor,  # noqa: F401
                      PitchEnergyPredictor, SpeechPredictor, Synthesizer, build_model)

__version__ = "0.1.0"
