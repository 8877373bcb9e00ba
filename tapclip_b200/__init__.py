"""tapclip_b200 — B200-native (sm_100a) drop-in for the hot path of 3300786/TAP-CLIP.

Public surface = the reference's: ``CLIPWrapper`` (models/clip_wrapper.py), ``FullModel`` (models/model_wrapper.py),
``PromptLearner``, ``AttributionMonitor``, ``PromptAdjustor``.  All compute goes through ``lib/libtapclip.so``
(C ABI in include/tapclip.h); importing the package is cheap, the library is loaded on first use and its absence
is an error (no CPU / PyTorch fallback).
"""
from .attribution_monitor import AttributionMonitor
from .clip_wrapper import CLIPWrapper
from .configs import get_model_config
from .eval_metrics import evaluate_accuracy, evaluate_per_class_accuracy
from .model_wrapper import FullModel
from .optim import FusedAdamW
from .preprocess import GpuPreprocess
from .prompt_adjustor import PromptAdjustor
from .prompt_learner import PromptLearner

__all__ = ["CLIPWrapper", "FullModel", "PromptLearner", "AttributionMonitor", "PromptAdjustor", "FusedAdamW",
           "GpuPreprocess", "get_model_config", "evaluate_accuracy", "evaluate_per_class_accuracy"]
__version__ = "0.1.0"
