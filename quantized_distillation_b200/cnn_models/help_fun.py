"""Per-batch loss + backward, evaluation and learning-rate schedules of the CNN
experiments (reference: cnn_models/help_fun.py).  No quantization happens here;
this is the code the hot path is sandwiched between."""
from __future__ import annotations

import math
import random

import torch
import torch.nn as nn
import torch.nn.functional as F


def _device_of(model) -> torch.device:
    return next(model.parameters()).device


def _to_device(batch, device):
    inputs, labels = batch
    return inputs.to(device, non_blocking=True), labels.to(device, non_blocking=True)


@torch.no_grad()
def evaluateModel(model, testLoader, fastEvaluation=True, maxExampleFastEvaluation=10000, k=1):
    """Top-k accuracy (reference: help_fun.py:30-58)."""
    model.eval()
    device = _device_of(model)
    correct = torch.zeros((), device=device)
    total = 0
    for batch in testLoader:
        inputs, labels = _to_device(batch, device)
        topk = model(inputs).topk(k, dim=1, largest=True, sorted=True)[1]
        correct += (topk == labels.view(-1, 1)).any(dim=1).sum()
        total += labels.numel()
        if fastEvaluation is True and total > maxExampleFastEvaluation:
            break
    return float(correct.item()) / max(total, 1)


def _teacher_mask(strategy, outputs, labels):
    """Which examples get the distillation term (reference: help_fun.py:97-111)."""
    name = strategy[0].lower()
    batch = outputs.size(0)
    if name == "always":
        return torch.ones(batch, dtype=torch.bool, device=outputs.device)
    if name == "incorrect_labels":
        return outputs.detach().argmax(dim=1) != labels
    if "entropy" in name:
        p = F.softmax(outputs.detach(), dim=1)
        entropy = -(p * torch.log2(p.clamp_min(1e-30))).sum(dim=1)
        if name == "cutoff_entropy":
            return entropy > strategy[1]
        if name == "random_entropy":
            draws = torch.tensor([random.random() for _ in range(batch)], device=outputs.device)
            return draws < entropy / math.log2(outputs.size(1))
    raise ValueError("ask_teacher_strategy is incorrectly formatted")


def distillation_loss(outputs, labels, teacher_outputs, temperature=2, weight_teacher_loss=0.7, criterion=None):
    """0.7 * T^2 * KL(softmax(t/T) || softmax(s/T)) + 0.3 * CE, with the reference's
    ``nn.KLDivLoss()`` default reduction = mean over ALL elements (help_fun.py:124-139)."""
    criterion = criterion or nn.CrossEntropyLoss()
    # explicit sum / numel: torch announces that reduction="mean" will change meaning
    kl = F.kl_div(F.log_softmax(outputs / temperature, dim=1), F.softmax(teacher_outputs / temperature, dim=1),
                  reduction="sum") / outputs.numel()
    return weight_teacher_loss * temperature ** 2 * kl + (1 - weight_teacher_loss) * criterion(outputs, labels)


def forward_and_backward(model, batch, idx_batch, epoch, criterion=None, use_distillation_loss=False, teacher_model=None,
                         temperature_distillation=2, ask_teacher_strategy="always", return_more_info=False,
                         return_tensor=False):
    """One student forward, teacher forward on the selected examples, loss, backward
    (reference: help_fun.py:60-158).  ``return_tensor=True`` keeps the loss on the
    device (no host sync per step); the default returns a Python float like the
    reference's ``loss.data[0]``."""
    if criterion is None:
        criterion = nn.CrossEntropyLoss()
    if use_distillation_loss is True and teacher_model is None:
        raise ValueError("To compute distillation loss you need to pass the teacher model")
    if not isinstance(ask_teacher_strategy, tuple):
        ask_teacher_strategy = (ask_teacher_strategy,)
    inputs, labels = _to_device(batch, _device_of(model))
    outputs = model(inputs)
    count_asked_teacher = 0
    if use_distillation_loss:
        mask = _teacher_mask(ask_teacher_strategy, outputs, labels)
        if ask_teacher_strategy[0].lower() == "always":
            with torch.no_grad():
                teacher_out = teacher_model(inputs)
            loss = distillation_loss(outputs, labels, teacher_out, temperature_distillation, criterion=criterion)
            count_asked_teacher = inputs.size(0)
        else:
            asked, rest = mask.nonzero().view(-1), (~mask).nonzero().view(-1)
            count_asked_teacher = int(asked.numel())
            loss = outputs.new_zeros(())
            if count_asked_teacher:
                with torch.no_grad():
                    teacher_out = teacher_model(inputs[asked])
                loss = loss + distillation_loss(outputs[asked], labels[asked], teacher_out, temperature_distillation,
                                                criterion=criterion)
            if rest.numel():
                loss = loss + criterion(outputs[rest], labels[rest])
    else:
        loss = criterion(outputs, labels)
    loss.backward()
    value = loss.detach() if return_tensor else float(loss.item())
    if return_more_info:
        return value, count_asked_teacher, inputs.size(0)
    return value


def add_gradient_noise(model, idx_batch, epoch, number_minibatches_per_epoch):
    """Annealed Gaussian gradient noise (reference: help_fun.py:160-170)."""
    std = (0.01 / (1 + epoch * number_minibatches_per_epoch + idx_batch) ** 0.55) ** 0.5
    for p in model.parameters():
        if p.grad is not None:
            p.grad.add_(torch.randn_like(p.grad) * std)


class LearningRateScheduler:
    """The reference's four schedules (help_fun.py:172-251): 'cifar100' step decay at
    epochs 60/120/160 by 0.2, 'imagenet' /10 every 30 epochs, and the validation-driven
    halving of 'generic' / 'quant_points_cifar100'."""

    def __init__(self, initial_learning_rate, learning_rate_type="generic"):
        if learning_rate_type not in ("generic", "cifar100", "imagenet", "quant_points_cifar100"):
            raise ValueError("Wrong learning rate type specified")
        self.initial_learning_rate = initial_learning_rate
        self.learning_rate_type = learning_rate_type
        self.current_learning_rate = initial_learning_rate
        self.old_validation_error = float("inf")
        self.epochs_since_validation_error_dropped = 0
        self.total_number_of_learning_rate_halves = 0
        self.epochs_to_wait_for_halving = 0
        self.best_validation_error = float("inf")

    def update_learning_rate(self, epoch, validation_error):
        kind = self.learning_rate_type
        if kind == "cifar100":
            power = 3 if epoch > 160 else 2 if epoch > 120 else 1 if epoch > 60 else 0
            self.current_learning_rate = self.initial_learning_rate * math.pow(0.2, power)
            return self.current_learning_rate, False
        if kind == "imagenet":
            return self.initial_learning_rate * (0.1 ** (epoch // 30)), False
        if kind == "generic":
            wait_reduce, wait_after_halving, wait_stop, max_halves = 10, 8, 30, 11
        else:
            wait_reduce, wait_after_halving, wait_stop, max_halves = 2, 0, float("inf"), float("inf")
        if validation_error + 0.001 < self.old_validation_error:          # 0.1% band
            self.old_validation_error = validation_error
            self.epochs_since_validation_error_dropped = 0
        else:
            self.epochs_since_validation_error_dropped += 1
        self.epochs_to_wait_for_halving = max(self.epochs_to_wait_for_halving - 1, 0)
        if self.epochs_since_validation_error_dropped >= wait_reduce and self.epochs_to_wait_for_halving == 0:
            self.epochs_to_wait_for_halving = wait_after_halving
            self.total_number_of_learning_rate_halves += 1
            self.current_learning_rate = self.current_learning_rate / 2
        stop = (self.epochs_since_validation_error_dropped > wait_stop
                or self.total_number_of_learning_rate_halves > max_halves)
        return self.current_learning_rate, stop


def synthetic_cifar_loader(num_batches, batch_size, seed=0, pin=True, num_classes=10):
    """CIFAR-shaped random data (3x32x32 float32, labels in [0, num_classes)): the
    datasets need network access; steps/s does not depend on pixel values."""
    g = torch.Generator().manual_seed(seed)
    batches = []
    for _ in range(num_batches):
        x = torch.randn(batch_size, 3, 32, 32, generator=g)
        y = torch.randint(0, num_classes, (batch_size,), generator=g)
        if pin and torch.cuda.is_available():
            x, y = x.pin_memory(), y.pin_memory()
        batches.append((x, y))
    return batches
