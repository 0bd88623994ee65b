"""fav-b200: B200-native corruption-sweep evaluation path for failure-aware vision.

Python host mirror of the reference-style objects (plain classes with reset(), dict results)
over the C ABI in include/fav_b200.h.  Import as ``import fav``.
"""
from .spec import CORRUPTIONS, CORRUPTION_ID, IMPLEMENTED, CorruptionConfig, MEAN_STD, SEVERITY, profile_for  # noqa: F401

__all__ = ["CORRUPTIONS", "CORRUPTION_ID", "IMPLEMENTED", "CorruptionConfig", "MEAN_STD", "SEVERITY",
           "profile_for", "VisionClassifier", "CorruptionSweep", "SweepConfig", "MetricsAccumulator",
           "UncertaintyGate", "TrustReplay"]


def __getattr__(name):        # torch-dependent classes load lazily so that `import fav` stays cheap
    if name == "VisionClassifier":
        from .classifier import VisionClassifier
        return VisionClassifier
    if name in ("CorruptionSweep", "SweepConfig", "MetricsAccumulator", "finalize", "partition"):
        from . import sweep
        return getattr(sweep, name)
    if name == "UncertaintyGate":
        from .gate import UncertaintyGate
        return UncertaintyGate
    if name == "TrustReplay":
        from .trust import TrustReplay
        return TrustReplay
    raise AttributeError(name)
