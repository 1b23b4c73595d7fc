"""Wide residual network for CIFAR (depth 6n+4, widening factor k), the student /
teacher family of BASELINE config 3 (reference: cnn_models/wide_resnet.py:28-89;
student WRN-16-22, teacher WRN-28-20, cifar10_wideResNet.py:65,91).

Module registration order matters: the training loop indexes
``model.parameters()`` to skip the first and last tensor
(``quantize_first_and_last_layer=False``), so the order here is the reference's:
stem conv, three stages of pre-activation blocks (bn1, conv1, bn2, conv2,
optional 1x1 shortcut), final bn, linear."""
from __future__ import annotations

import math

import torch.nn as nn
import torch.nn.functional as F


class wide_basic(nn.Module):
    """Pre-activation block: BN-ReLU-conv3x3-dropout-BN-ReLU-conv3x3(stride) + shortcut."""

    def __init__(self, in_planes, planes, dropout_rate, stride=1):
        super().__init__()
        self.bn1 = nn.BatchNorm2d(in_planes)
        self.conv1 = nn.Conv2d(in_planes, planes, kernel_size=3, padding=1, bias=True)
        self.dropout = nn.Dropout(p=dropout_rate)
        self.bn2 = nn.BatchNorm2d(planes)
        self.conv2 = nn.Conv2d(planes, planes, kernel_size=3, stride=stride, padding=1, bias=True)
        self.shortcut = nn.Sequential()
        if stride != 1 or in_planes != planes:
            self.shortcut = nn.Sequential(nn.Conv2d(in_planes, planes, kernel_size=1, stride=stride, bias=True))

    def forward(self, x):
        y = self.dropout(self.conv1(F.relu(self.bn1(x))))
        y = self.conv2(F.relu(self.bn2(y)))
        return y + self.shortcut(x)


def _init(module):
    if isinstance(module, nn.Conv2d):
        nn.init.xavier_uniform_(module.weight, gain=math.sqrt(2))
        nn.init.constant_(module.bias, 0)
    elif isinstance(module, nn.BatchNorm2d):
        nn.init.constant_(module.weight, 1)
        nn.init.constant_(module.bias, 0)


class Wide_ResNet(nn.Module):
    def __init__(self, depth, widen_factor, dropout_rate, num_classes):
        super().__init__()
        if (depth - 4) % 6 != 0:
            raise ValueError("Wide-resnet depth should be 6n+4")
        n, k = (depth - 4) // 6, widen_factor
        widths = [16, 16 * k, 32 * k, 64 * k]
        self.in_planes = 16
        self.conv1 = nn.Conv2d(3, widths[0], kernel_size=3, stride=1, padding=1, bias=True)
        self.layer1 = self._stage(widths[1], n, dropout_rate, stride=1)
        self.layer2 = self._stage(widths[2], n, dropout_rate, stride=2)
        self.layer3 = self._stage(widths[3], n, dropout_rate, stride=2)
        self.bn1 = nn.BatchNorm2d(widths[3], momentum=0.9)
        self.linear = nn.Linear(widths[3], num_classes)
        self.apply(_init)

    def _stage(self, planes, blocks, dropout_rate, stride):
        layers = []
        for s in [stride] + [1] * (blocks - 1):
            layers.append(wide_basic(self.in_planes, planes, dropout_rate, s))
            self.in_planes = planes
        return nn.Sequential(*layers)

    def forward(self, x):
        y = self.layer3(self.layer2(self.layer1(self.conv1(x))))
        y = F.relu(self.bn1(y))
        y = F.avg_pool2d(y, 8)
        return self.linear(y.view(y.size(0), -1))
