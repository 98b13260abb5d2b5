"""gemlite_b200 -- Blackwell (sm_100a) native low-bit fused dequant + matmul, GemLite-compatible API.

    from gemlite_b200 import GemLiteLinear, DType, helper

The public names mirror `gemlite/__init__.py:5-22` of the reference so that `import gemlite_b200 as
gemlite` is a drop-in for the GemLiteLinear / helper surface.
"""
__version__ = "0.1.0"
__author__ = "gemlite_b200 authors"
__credits__ = "API modelled on mobiusml/gemlite"

from .dtypes import DType  # noqa: F401
from .core import (  # noqa: F401
    GemLiteLinear, GemLiteLinearTriton, GEMLITE_ACC_DTYPE, GEMLITE_MATMUL_TYPES, GEMLITE_MATMUL_TYPES_MAPPING,
    forward_functional, get_matmul_type, get_default_gemv,
    set_autotune, set_autotune_setting, set_packing_bitwidth, set_acc_dtype, set_kernel_caching,
)
from . import helper  # noqa: F401

load_config = GemLiteLinear.load_config
cache_config = GemLiteLinear.cache_config
reset_config = GemLiteLinear.reset_config
