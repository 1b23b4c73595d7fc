"""Drop-in for the reference's ``quantization`` package
(quantization/__init__.py:3-8): same names, CUDA kernels behind them."""
import torch

USE_CUDA = torch.cuda.is_available()
from .quant_functions import (ScalingFunction, SearchSorted, nonUniformQuantization,  # noqa: E402
                              nonUniformQuantization_variable, uniformQuantization, uniformQuantization_variable)
from . import help_functions  # noqa: E402,F401

__all__ = ("uniformQuantization", "ScalingFunction", "nonUniformQuantization", "uniformQuantization_variable",
           "nonUniformQuantization_variable")
