"""covidcurve_b200: B200-native (sm_100a, FP64 CUDA) drop-in for COVIDCurve's batched IFR log-likelihood path."""
from .spec import HotPathSpec  # noqa: F401

__version__ = "0.1.0"
