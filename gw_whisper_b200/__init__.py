"""gw_whisper_b200 -- B200-native (sm_100a) implementation of GW-Whisper's sliding-window inference
path behind the reference's own Python signatures.  See DESIGN.md / INTEGRATION.md."""
from .encoder import B200WhisperEncoder, WhisperGeometry, load_dora_adapter  # noqa: F401
from .frontend import logmel_features, resample_timeseries, LogMelFeatureExtractor  # noqa: F401
from .models import (  # noqa: F401
    two_channel_ligo_binary_classifier,
    one_channel_ligo_binary_classifier,
    glitch_one_channel_classifier,
    replace_softmax_by_mutual_subtraction,
)
from . import evaluate  # noqa: F401  (FAR / sensitive distance of a trigger list: MLGWSC-1/evaluate.py get_stats)
from .qfrontend import (  # noqa: F401
    QScanB200, QTransformAdapter, TrainQTransformAdapter, GWWhisperClassifier, remove_softmax_from_classifier)

__all__ = [
    "QScanB200", "QTransformAdapter", "TrainQTransformAdapter", "GWWhisperClassifier", "remove_softmax_from_classifier",
    "B200WhisperEncoder", "WhisperGeometry", "load_dora_adapter", "logmel_features",
    "resample_timeseries", "LogMelFeatureExtractor", "two_channel_ligo_binary_classifier",
    "one_channel_ligo_binary_classifier", "glitch_one_channel_classifier", "replace_softmax_by_mutual_subtraction",
]
