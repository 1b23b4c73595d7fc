"""Generates tests/golden/reference_models.npz: the reference's two CIFAR model definitions
(cnn_models/conv_forward_model.py:42-163 ConvolForwardNet, cnn_models/wide_resnet.py:50-89 Wide_ResNet), built
unmodified from /root/reference in SMALL configurations, with their parameter / state-dict order, their weights, one
input batch and the logits they produce in eval() and in train() mode (no dropout: deterministic).  The hot path
quantizes ``model.parameters()`` in registration order and ``quantize_first_and_last_layer=False`` skips the first and
the last entry of that list (conv_forward_model.py:237-239), so the order is part of the contract.

Run in the build container only:  python tests/golden/make_golden_models.py
"""
import os
import sys
import warnings

import numpy as np
import torch

REF = "/root/reference"
sys.path.insert(0, REF)
warnings.filterwarnings("ignore")
import cnn_models.conv_forward_model as RC  # noqa: E402
import cnn_models.wide_resnet as RW  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "reference_models.npz")

SMALL_CONV = {"spec_conv_layers": [(6, 3, 3), (8, 5, 5), (8, 3, 3)], "spec_max_pooling": [(0, 2, 2), (2, 2, 2)],
              "spec_dropout_rates": [], "spec_linear": [24, 12], "width": 16, "height": 16}


def dump(store, tag, model, x):
    store[tag + "_param_names"] = np.array([n for n, _ in model.named_parameters()])
    store[tag + "_state_names"] = np.array(list(model.state_dict().keys()))
    for k, v in model.state_dict().items():
        store[f"{tag}_sd_{k}"] = v.detach().numpy().copy()
    store[tag + "_x"] = x.numpy().copy()
    model.eval()
    with torch.no_grad():
        store[tag + "_y_eval"] = model(x).numpy().copy()
    model.train()
    with torch.no_grad():
        store[tag + "_y_train"] = model(x).numpy().copy()             # batch statistics; also moves the running stats
    for k, v in model.state_dict().items():
        if "running" in k or "num_batches" in k:
            store[f"{tag}_after_{k}"] = v.detach().numpy().copy()


def main():
    store = {}
    torch.manual_seed(5)
    for tag, bn, affine in (("conv_bn_affine", True, True), ("conv_bn", True, False), ("conv_plain", False, False)):
        m = RC.ConvolForwardNet(**SMALL_CONV, useBatchNorm=bn, useAffineTransformInBatchNorm=affine)
        with torch.no_grad():                                        # non-trivial batch-norm state
            for k, v in m.state_dict().items():
                if k.endswith("running_mean"):
                    v.normal_(0, 0.2)
                elif k.endswith("running_var"):
                    v.uniform_(0.5, 1.5)
        dump(store, tag, m, torch.randn(5, 3, 16, 16))
    m = RW.Wide_ResNet(depth=10, widen_factor=1, dropout_rate=0.0, num_classes=10)
    dump(store, "wrn_10_1", m, torch.randn(4, 3, 32, 32))
    # the paper specifications themselves (conv_forward_model.py:30-40)
    store["teacherModelSpec"] = np.array(repr(RC.teacherModelSpec))
    store["smallerModelSpec"] = np.array(repr(RC.smallerModelSpec))
    np.savez_compressed(OUT, **store)
    print("wrote", OUT, os.path.getsize(OUT) // 1024, "KiB; torch", torch.__version__)


if __name__ == "__main__":
    main()
