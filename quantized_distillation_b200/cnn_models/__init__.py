"""Callers of the hot path ("next" rows / harness): the CIFAR
models and the quantized-distillation / differentiable-quantization loops of the
reference's ``cnn_models`` package, re-hosted on current PyTorch with the
per-step quantization choreography replaced by the multi-tensor CUDA plan."""
from . import conv_forward_model, help_fun, wide_resnet  # noqa: F401
from .conv_forward_model import (ConvolForwardNet, optimize_quantization_points, smallerModelSpec, teacherModelSpec,  # noqa: F401
                                 train_model, train_model_quantized)
from .wide_resnet import Wide_ResNet  # noqa: F401
