"""
rajepy_b200 -- B200-native engine for RaJePy's hot path (jet-grid fill and line-of-sight
radiative transfer) behind the reference's JetModel API.  See DESIGN.md.
"""
from . import hostmath  # noqa: F401
from . import logger  # noqa: F401
from .jetmodel import (JetModel, check_model_params, flux_ff_time_series,  # noqa: F401
                       reorder_axes)
from ._cabi import EngineError  # noqa: F401

__version__ = "0.1.0"
