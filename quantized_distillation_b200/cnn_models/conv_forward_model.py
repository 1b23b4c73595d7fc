"""CIFAR conv-net and the two training loops that call the quantization hot path
on every step (reference: cnn_models/conv_forward_model.py):

* ``train_model(quantizeWeights=True)`` -- quantized distillation (:165-393);
  ``train_model_quantized`` is the alias BASELINE.json's north_star names.
* ``optimize_quantization_points`` -- differentiable quantization of the
  centroids (:395-592).

Same keyword arguments and return values as the reference.  What changes is the
per-step choreography around the model forward/backward: the reference saves
``state_dict()``, rebinds every ``p.data`` to a freshly quantized tensor (about
12 launches per tensor) and copies the weights back with ``load_state_dict``;
here a :class:`QuantizationPlan` snapshots, quantizes IN PLACE and restores all
tensors with one launch each, which also keeps ``DistributedDataParallel``'s
parameter references valid.  Unlike the reference (:369-374) exceptions are not
swallowed.
"""
from __future__ import annotations

import copy
import time

import torch
import torch.nn as nn
import torch.nn.functional as F
import torch.optim as optim

from .. import quantization
from ..plan import CentroidPlan, QuantizationPlan
from . import help_fun as cnn_hf

# paper specifications (reference :30-40): teacher ~5.3 M parameters, student ~1 M
teacherModelSpec = {"spec_conv_layers": [(76, 3, 3), (76, 3, 3), (126, 3, 3), (126, 3, 3), (148, 3, 3), (148, 3, 3),
                                         (148, 3, 3), (148, 3, 3)],
                    "spec_max_pooling": [(1, 2, 2), (3, 2, 2), (7, 2, 2)],
                    "spec_dropout_rates": [(1, 0.2), (3, 0.3), (7, 0.35), (8, 0.4), (9, 0.4)],
                    "spec_linear": [1200, 1200], "width": 32, "height": 32}
smallerModelSpec = {"spec_conv_layers": [(75, 5, 5), (50, 5, 5), (50, 5, 5), (25, 5, 5)],
                    "spec_max_pooling": [(1, 2, 2), (3, 2, 2)],
                    "spec_dropout_rates": [(1, 0.2), (3, 0.3), (4, 0.4)],
                    "spec_linear": [500], "width": 32, "height": 32}


class ConvolForwardNet(nn.Module):
    """Stack of same-padded conv layers with max-pooling / dropout inserted at given
    positions, then linear layers and a 10-way output layer, ReLU everywhere
    (reference :42-163).

    The output layer is registered BEFORE the layer lists on purpose: that is the
    reference's registration order (:124-132), so ``parameters()[0]`` is
    ``out_layer.weight`` and ``quantize_first_and_last_layer=False`` skips the
    same two tensors as in the reference."""

    def __init__(self, width, height, spec_conv_layers, spec_max_pooling, spec_linear, spec_dropout_rates,
                 useBatchNorm=False, useAffineTransformInBatchNorm=False):
        super().__init__()
        self.width, self.height = width, height
        self.useBatchNorm = useBatchNorm
        convs, norms, pools, drops, linears = [], [], [], [], []
        channels = 3
        for filters, kh, kw in spec_conv_layers:
            conv = nn.Conv2d(channels, filters, kernel_size=(kh, kw), padding=((kh - 1) // 2, (kw - 1) // 2))
            nn.init.xavier_uniform_(conv.weight, nn.init.calculate_gain("conv2d"))
            convs.append(conv)
            norms.append(nn.BatchNorm2d(filters, affine=useAffineTransformInBatchNorm))
            channels = filters
        self.max_pooling_positions = [pos for pos, _, _ in spec_max_pooling]
        pools = [nn.MaxPool2d((kh, kw)) for _, kh, kw in spec_max_pooling]
        self.dropout_positions = [pos for pos, _ in spec_dropout_rates]
        drops = [nn.Dropout2d(rate) if pos < len(convs) else nn.Dropout(rate) for pos, rate in spec_dropout_rates]
        features = channels * width * height // 2 ** (2 * len(pools))
        for units in spec_linear:
            lin = nn.Linear(features, units)
            nn.init.xavier_uniform_(lin.weight, nn.init.calculate_gain("linear"))
            linears.append(lin)
            norms.append(nn.BatchNorm1d(units, affine=useAffineTransformInBatchNorm))
            features = units
        self.out_layer = nn.Linear(features, 10)
        nn.init.xavier_uniform_(self.out_layer.weight, nn.init.calculate_gain("linear"))
        self.conv_layers = nn.ModuleList(convs)
        self.max_pooling_layers = nn.ModuleList(pools)
        self.dropout_layers = nn.ModuleList(drops)
        self.linear_layers = nn.ModuleList(linears)
        self.batchNormalizationLayers = nn.ModuleList(norms)
        self.num_conv_layers = len(convs)
        self.total_num_layers = len(convs) + len(linears)

    def forward(self, x):
        for i in range(self.total_num_layers):
            if i < self.num_conv_layers:
                x = F.relu(self.conv_layers[i](x))
            else:
                if i == self.num_conv_layers:
                    x = x.view(x.size(0), -1)
                x = F.relu(self.linear_layers[i - self.num_conv_layers](x))
            if self.useBatchNorm:
                x = self.batchNormalizationLayers[i](x)
            if i in self.max_pooling_positions:
                x = self.max_pooling_layers[self.max_pooling_positions.index(i)](x)
            if i in self.dropout_positions:
                x = self.dropout_layers[self.dropout_positions.index(i)](x)
        return F.relu(self.out_layer(x))


# ------------------------------------------------------------------------------------------
# quantized distillation
# ------------------------------------------------------------------------------------------
class _GraphedStep:
    """One whole training step captured in a CUDA graph: static input buffers, one replay per step."""

    def __init__(self, step_fn, example_batch, device, stream, optimizer, static_grads=False, collective=False):
        self.ok = False
        try:
            x, y = example_batch
            self.x = torch.empty(x.shape, dtype=x.dtype, device=device)
            self.y = torch.empty(y.shape, dtype=y.dtype, device=device)
            self.x.copy_(x, non_blocking=True)
            self.y.copy_(y, non_blocking=True)
            torch.cuda.synchronize(device)
            if not static_grads:
                optimizer.zero_grad(set_to_none=True)        # gradients are re-created inside the graph's pool
            # (static_grads: they are views of FlatDataParallel's flat buffer, allocated once, address-stable)
            self.graph = torch.cuda.CUDAGraph()
            # a captured step that holds an NCCL collective: other threads of the process (the NCCL
            # watchdog) may touch the CUDA API meanwhile, which only thread-local capture tolerates
            mode = {"capture_error_mode": "thread_local"} if collective else {}
            with torch.cuda.graph(self.graph, stream=stream, **mode):
                self.loss, self.asked, self.total = step_fn((self.x, self.y))
            self.ok = True
        except Exception as e:                               # pragma: no cover - depends on driver / torch build
            import warnings
            warnings.warn(f"CUDA graph capture of the training step failed, running eagerly: {e}")
            try:
                torch.cuda.synchronize(device)
            except Exception:
                pass

    def matches(self, batch):
        return batch[0].shape == self.x.shape and batch[1].shape == self.y.shape

    def run(self, batch):
        self.x.copy_(batch[0], non_blocking=True)
        self.y.copy_(batch[1], non_blocking=True)
        self.graph.replay()
        return self.loss, self.asked, self.total



def _selected_parameters(model, quantize_first_and_last_layer):
    params = list(model.parameters())
    if quantize_first_and_last_layer is False:
        params = params[1:-1]                                             # reference :237-239
    return params


def _uniform_levels(quantizationFunctionToUse, numBits):
    name = quantizationFunctionToUse.lower()
    if name == "uniformAbsMaxScaling".lower():
        return 2 ** (numBits - 1), "absmax"                               # reference :206-208 (broken scaling there)
    if name == "uniformLinearScaling".lower():
        return 2 ** numBits, "linear"                                     # reference :209-211
    raise ValueError("The specified quantization function is not present")


class WeightQuantizer:
    """``quantize_weights_model`` / ``backward_quant_weights_model`` of the reference
    (closures at :235-266) lifted to an object that owns the multi-tensor plan."""

    def __init__(self, model, numBits, bucket_size, quantizationFunctionToUse="uniformLinearScaling",
                 backprop_quantization_style="none", quantize_first_and_last_layer=True, *, stochastic_rounding=False,
                 max_element=False, subtract_mean=False):
        style = "none" if backprop_quantization_style is None else backprop_quantization_style.lower()
        if style not in ("none", "truncated", "complicated"):
            raise ValueError("The specified backprop_quantization_style not recognized")
        self.style = style
        self.s, self.scaling = _uniform_levels(quantizationFunctionToUse, numBits)
        self.bucket_size = bucket_size
        self.params = _selected_parameters(model, quantize_first_and_last_layer)
        # the options only the NMT loop passes (translation_models/model.py:162-164, 198-204: stochasticRounding,
        # maxElementAllowedForQuantization, subtractMeanInQuantization); the multi-tensor plan is the plain
        # deterministic op, so any of them selects the per-tensor fused kernel, which implements all three
        self.options = {"stochastic_rounding": bool(stochastic_rounding), "max_element": max_element,
                        "subtract_mean": bool(subtract_mean)}
        extra = self.options["stochastic_rounding"] or max_element is not False or self.options["subtract_mean"]
        if extra and self.scaling == "linear":
            if style == "complicated":
                # reference: backward raises for subtract_mean (quant_functions.py:329-330) and re-runs its forward with
                # max_element / stochastic rounding (:341-347), which the backward kernel does not take
                raise NotImplementedError("backprop_quantization_style 'complicated' is not available together with "
                                          "stochastic_rounding / max_element / subtract_mean")
            self.plan = None
            self._master = [torch.empty_like(p.data) for p in self.params]
            return
        if self.scaling != "linear":
            # 'uniformAbsMaxScaling' (:206-208) cannot execute in the reference; the intended semantics are an opt-in
            # extension without a parity target (quantization.quant_functions.ALLOW_UNPINNED_SCALING), per tensor
            if not quantization.quant_functions.ALLOW_UNPINNED_SCALING:
                raise NotImplementedError("absmax scaling does not execute in the reference (quant_functions.py:119-126); "
                                          "set quantization.quant_functions.ALLOW_UNPINNED_SCALING = True for the extension")
            if style == "complicated":
                raise ValueError("Linear scaling is necessary to backpropagate")              # quant_functions.py:326-327
            self.plan = None
            self._master = [torch.empty_like(p.data) for p in self.params]
            return
        self.plan = QuantizationPlan(self.params, self.s, bucket_size)

    def quantize_weights_model(self, save=True):
        """fp32 weights -> shadow buffer, then every tensor quantized in place."""
        if self.style == "truncated":
            torch._foreach_clamp_min_([p.data for p in self.params], -1.0)   # p.data.clamp_(-1, 1), reference :240-241
            torch._foreach_clamp_max_([p.data for p in self.params], 1.0)
        if self.plan is None:                                  # NMT-loop options / absmax extension: one fused launch per tensor
            if save:
                torch._foreach_copy_(self._master, [p.data for p in self.params])
            for p in self.params:
                quantization.uniformQuantization(p.data, self.s, type_of_scaling=self.scaling, bucket_size=self.bucket_size,
                                                 modify_in_place=True, **self.options)
            return
        if save:
            self.plan.save_and_quantize_()          # shadow copy + in-place quantization, one launch
        else:
            self.plan.quantize_()

    def restore_weights_model(self):
        if self.plan is None:
            torch._foreach_copy_([p.data for p in self.params], self._master)
            return
        self.plan.restore_master()

    def backward_quant_weights_model(self):
        if self.style == "none":
            return
        if self.plan is None:                                                 # 'truncated' is scaling-agnostic (:263-264)
            for p in self.params:
                p.grad.data.masked_fill_(p.data.abs() > 1, 0.0)
            return
        grads = []
        for p in self.params:
            if p.grad is None:
                raise ValueError("backward_quant_weights_model needs gradients on every quantized parameter")
            grads.append(p.grad.data if p.grad.is_contiguous() else p.grad.data.contiguous())
        self.plan.backward_(grads, self.style)
        for p, g in zip(self.params, grads):
            if g.data_ptr() != p.grad.data_ptr():
                p.grad.data.copy_(g)


def quantize_weights_model(model, numBits, bucket_size=None, quantize_first_and_last_layer=True):
    """One-shot post-training quantization of a model's weights, in place
    (what the drivers do tensor by tensor, cifar10_test.py:312-317)."""
    WeightQuantizer(model, numBits, bucket_size, quantize_first_and_last_layer=quantize_first_and_last_layer) \
        .quantize_weights_model(save=False)
    return model


def train_model(model, train_loader, test_loader, initial_learning_rate=0.001, use_nesterov=True,
                initial_momentum=0.9, weight_decayL2=0.00022, epochs_to_train=100, print_every=500,
                learning_rate_style="generic", use_distillation_loss=False, teacher_model=None,
                quantizeWeights=False, numBits=8, grad_clipping_threshold=False, start_epoch=0,
                bucket_size=None, quantizationFunctionToUse="uniformLinearScaling",
                backprop_quantization_style="none", estimate_quant_grad_every=1, add_gradient_noise=False,
                ask_teacher_strategy=("always", None), quantize_first_and_last_layer=True,
                mix_with_differentiable_quantization=False, *, max_steps=None, verbose=True, evaluate=True,
                step_hook=None, cuda_graph_step=False, fused_optimizer_step=False):
    """SGD training with optional distillation loss and optional per-step weight
    quantization (reference :165-393; same positional/keyword arguments, the
    keyword-only ones after ``*`` are additions).

    ``cuda_graph_step=True`` (single process, ``ask_teacher_strategy`` 'always',
    ``estimate_quant_grad_every`` 1): after three eager steps the whole step --
    save+quantize, student/teacher forward, backward, restore, gradient fix-up,
    SGD update -- is captured once in a CUDA graph and replayed; each step then
    costs one H2D copy of the batch and one graph launch.  Same arithmetic, same
    kernels; the graph is re-captured when the learning rate changes.

    ``fused_optimizer_step=True`` (quantized training, bucket of at most 512, no gradient noise /
    clipping): restore, gradient fix-up, the SGD update and the NEXT step's save-and-quantize become
    one kernel over the quantized tensors (``QuantizationPlan.fused_step_``, 24 instead of 52 bytes
    per weight); the live parameters then always hold the quantized weights and the full-precision
    ones live in the plan's master buffer.  Same arithmetic as ``torch.optim.SGD`` + the unfused ops."""
    if use_distillation_loss is True and teacher_model is None:
        raise ValueError("To compute distillation loss you have to pass the teacher model")
    if teacher_model is not None:
        teacher_model.eval()
    learning_rate_style = learning_rate_style.lower()
    lr_scheduler = cnn_hf.LearningRateScheduler(initial_learning_rate, learning_rate_style)
    new_learning_rate = initial_learning_rate
    optimizer = optim.SGD(model.parameters(), lr=initial_learning_rate, nesterov=use_nesterov, momentum=initial_momentum,
                          weight_decay=weight_decayL2)
    start_time = time.time()
    pred_accuracy_epochs, percentages_asked_teacher, losses_epochs = [], [], []
    informationDict = {}
    last_loss_saved = float("inf")
    steps_since_estimate = 1
    batches_per_epoch = len(train_loader)
    quantizer = None
    if quantizeWeights:
        quantizer = WeightQuantizer(model, numBits, bucket_size, quantizationFunctionToUse, backprop_quantization_style,
                                    quantize_first_and_last_layer)
    if print_every > batches_per_epoch:
        print_every = max(batches_per_epoch // 2, 1)
    total_steps = 0
    stop = False
    epoch = start_epoch
    device = cnn_hf._device_of(model)
    state = {"since": steps_since_estimate}
    # data parallelism: FlatDataParallel exposes reduce_gradients(); DDP reduces inside backward
    reduce_gradients = getattr(model, "reduce_gradients", None)
    flat_dp = reduce_gradients is not None

    fused = bool(fused_optimizer_step and quantizer is not None and device.type == "cuda" and estimate_quant_grad_every == 1
                 and not add_gradient_noise and grad_clipping_threshold is False
                 and bucket_size is not None and bucket_size <= 512 and quantizer.plan is not None)
    rest_optimizer = None
    if fused:
        chosen = {id(p) for p in quantizer.params}
        rest = [p for p in model.parameters() if id(p) not in chosen]      # e.g. first / last layer left unquantized
        if rest:
            rest_optimizer = optim.SGD(rest, lr=initial_learning_rate, nesterov=use_nesterov, momentum=initial_momentum,
                                       weight_decay=weight_decayL2)
        state["lr"] = initial_learning_rate
        quantizer.quantize_weights_model()                                 # master <- weights, live <- quantized, once

    def fused_step(data, idx_minibatch=1, epoch=0):
        """The same step with the tail fused: live parameters are already quantized on entry."""
        model.zero_grad(set_to_none=False)
        loss, c_teach, c_total = cnn_hf.forward_and_backward(
            model, data, idx_minibatch, epoch, use_distillation_loss=use_distillation_loss, teacher_model=teacher_model,
            ask_teacher_strategy=ask_teacher_strategy, return_more_info=True, return_tensor=True)
        if reduce_gradients is not None:
            reduce_gradients()
        grads = []
        for p in quantizer.params:
            if p.grad is None:
                raise ValueError("every quantized parameter needs a gradient")
            if not p.grad.is_contiguous():
                p.grad = p.grad.contiguous()
            grads.append(p.grad.data)
        quantizer.plan.fused_step_(grads, quantizer.style, state["lr"], initial_momentum, weight_decayL2, use_nesterov)
        if rest_optimizer is not None:
            rest_optimizer.step()
        return loss, c_teach, c_total

    def one_step(data, idx_minibatch=1, epoch=0):
        """One training step of the reference loop (:280-322) on one batch."""
        if fused:
            return fused_step(data, idx_minibatch, epoch)
        quantize_now = quantizer is not None and state["since"] >= estimate_quant_grad_every
        if quantize_now:
            quantizer.quantize_weights_model()                            # :286-287
        model.zero_grad(set_to_none=False)
        loss, c_teach, c_total = cnn_hf.forward_and_backward(
            model, data, idx_minibatch, epoch, use_distillation_loss=use_distillation_loss, teacher_model=teacher_model,
            ask_teacher_strategy=ask_teacher_strategy, return_more_info=True, return_tensor=True)
        if reduce_gradients is not None:
            reduce_gradients()                                            # ONE all-reduce of the flat gradient buffer
        if quantize_now:
            quantizer.restore_weights_model()                             # :302
        if add_gradient_noise and not quantizeWeights:
            cnn_hf.add_gradient_noise(model, idx_minibatch, epoch, batches_per_epoch)
        if grad_clipping_threshold is not False:
            for p in model.parameters():
                if p.grad is not None:
                    p.grad.clamp_(-grad_clipping_threshold, grad_clipping_threshold)
        if quantize_now:
            quantizer.backward_quant_weights_model()                      # :315
        optimizer.step()
        if state["since"] >= estimate_quant_grad_every:
            state["since"] = 0
        state["since"] += 1
        return loss, c_teach, c_total

    strategy_name = (ask_teacher_strategy[0] if isinstance(ask_teacher_strategy, tuple) else ask_teacher_strategy).lower()
    multi = torch.distributed.is_available() and torch.distributed.is_initialized() and torch.distributed.get_world_size() > 1
    # stock DDP drives its reducer from autograd hooks on the host and cannot be replayed; the flat
    # wrapper's single all-reduce can, so a captured step is available for any world size with it
    graph_ok = bool(cuda_graph_step and device.type == "cuda" and estimate_quant_grad_every == 1 and not add_gradient_noise
                    and strategy_name == "always" and (not multi or flat_dp))
    side_stream = torch.cuda.Stream(device) if graph_ok else None
    graphed = None
    try:
        for epoch in range(start_epoch, epochs_to_train + start_epoch):
            model.train()
            running = torch.zeros((), device=cnn_hf._device_of(model))
            asked, seen = 0, 0
            for idx_minibatch, data in enumerate(train_loader, start=1):
                if graph_ok and graphed is None and total_steps >= 3:
                    graphed = _GraphedStep(one_step, data, device, side_stream, optimizer, static_grads=flat_dp,
                                           collective=flat_dp and multi)
                    captured = graphed.ok
                    if multi:                                      # replay a collective only if EVERY rank captured it
                        flag = torch.tensor([1 if captured else 0], dtype=torch.int32, device=device)
                        torch.distributed.all_reduce(flag, op=torch.distributed.ReduceOp.MIN)
                        captured = bool(flag.item())
                    informationDict["cuda_graph_step"] = captured
                    if not captured:
                        graph_ok, graphed = False, None
                if graphed is not None and graphed.matches(data):
                    loss, c_teach, c_total = graphed.run(data)
                elif graph_ok:
                    with torch.cuda.stream(side_stream):           # warm-up steps run where the capture will
                        loss, c_teach, c_total = one_step(data, idx_minibatch, epoch)
                    torch.cuda.current_stream(device).wait_stream(side_stream)
                else:
                    loss, c_teach, c_total = one_step(data, idx_minibatch, epoch)
                asked += c_teach
                seen += c_total
                running += loss
                total_steps += 1
                if step_hook is not None:
                    step_hook(total_steps, loss)
                if idx_minibatch % print_every == 0:
                    last_loss_saved = float(running.item()) / print_every
                    running.zero_()
                    if verbose:
                        msg = "Time Elapsed: {:.1f}s, [Start Epoch: {}, Epoch: {}, Minibatch: {}], loss: {:3f}".format(
                            time.time() - start_time, start_epoch + 1, epoch + 1, idx_minibatch, last_loss_saved)
                        if pred_accuracy_epochs:
                            msg += " Last prediction accuracy: {:2f}%".format(pred_accuracy_epochs[-1] * 100)
                        print(msg)
                if max_steps is not None and total_steps >= max_steps:
                    stop = True
                    break
            percentages_asked_teacher.append(asked / seen if seen else 0)
            losses_epochs.append(last_loss_saved)
            if evaluate:
                pred_accuracy_epochs.append(cnn_hf.evaluateModel(model, test_loader, fastEvaluation=False))
                if verbose:
                    print(" === Epoch: {} - prediction accuracy {:2f}% === ".format(epoch + 1, pred_accuracy_epochs[-1] * 100))
            if stop:
                break
            if mix_with_differentiable_quantization and epoch != start_epoch + epochs_to_train - 1:
                # the differentiable step works on a copy and hands back its state dict (reference :342-353)
                quantized_state_dict = optimize_quantization_points(
                    model, train_loader, test_loader, new_learning_rate, initial_momentum=initial_momentum,
                    epochs_to_train=1, print_every=print_every, use_nesterov=use_nesterov,
                    learning_rate_style=learning_rate_style, numPointsPerTensor=2 ** numBits,
                    assignBitsAutomatically=True, bucket_size=bucket_size, use_distillation_loss=True,
                    initialize_method="quantiles", quantize_first_and_last_layer=quantize_first_and_last_layer,
                    verbose=verbose, evaluate=evaluate, max_steps=max_steps)[0]
                model.load_state_dict(quantized_state_dict)
                if fused:
                    quantizer.quantize_weights_model()                     # master <- the loaded weights, live <- quantized
                losses_epochs.append(last_loss_saved)
                if evaluate:
                    pred_accuracy_epochs.append(cnn_hf.evaluateModel(model, test_loader, fastEvaluation=False))
            error = 1 - pred_accuracy_epochs[-1] if pred_accuracy_epochs else 1.0
            new_learning_rate, stop_training = lr_scheduler.update_learning_rate(epoch, error)
            if stop_training is True:
                break
            for opt in (optimizer, rest_optimizer):
                for group in (opt.param_groups if opt is not None else []):
                    if group["lr"] != new_learning_rate:
                        graphed = None                                     # the captured SGD update holds the old rate
                    group["lr"] = new_learning_rate
            state["lr"] = new_learning_rate
    except KeyboardInterrupt:
        informationDict["errorFlag"] = False
        informationDict["numEpochsTrained"] = epoch - start_epoch
    else:
        informationDict["errorFlag"] = False
        informationDict["numEpochsTrained"] = epoch + 1 - start_epoch
    if quantizer is not None and not fused:                                # (fused: the live weights already are)
        quantizer.quantize_weights_model(save=False)                       # final weights are returned quantized (:384-385)
    informationDict["fused_optimizer_step"] = fused
    if mix_with_differentiable_quantization:
        informationDict["numEpochsTrained"] *= 2
    informationDict["percentages_asked_teacher"] = percentages_asked_teacher
    informationDict["predictionAccuracy"] = pred_accuracy_epochs
    informationDict["lossSaved"] = losses_epochs
    informationDict["numStepsTrained"] = total_steps
    return model, informationDict


def train_model_quantized(model, train_loader, test_loader, numBits=8, bucket_size=None, **kwargs):
    """``train_model(..., quantizeWeights=True)``: the entry point BASELINE.json's
    north_star names (the reference spells it through the ``quantizeWeights`` flag)."""
    return train_model(model, train_loader, test_loader, quantizeWeights=True, numBits=numBits, bucket_size=bucket_size,
                       **kwargs)


# ------------------------------------------------------------------------------------------
# differentiable quantization
# ------------------------------------------------------------------------------------------
def _capture_point_graphs(quantize_all, point_gradients, device):
    """Captures the two per-step launch sequences of the differentiable-quantization loop.
    Returns (forward_graph, backward_graph, gradient tensors) or False when capture fails."""
    try:
        torch.cuda.synchronize(device)
        side = torch.cuda.Stream(device)
        side.wait_stream(torch.cuda.current_stream(device))
        with torch.cuda.stream(side):                       # warm the capture stream's workspace / caches
            quantize_all()
            point_gradients()
        torch.cuda.current_stream(device).wait_stream(side)
        torch.cuda.synchronize(device)
        g_fwd, g_bwd = torch.cuda.CUDAGraph(), torch.cuda.CUDAGraph()
        with torch.cuda.graph(g_fwd):
            quantize_all()
        with torch.cuda.graph(g_bwd):
            grads = point_gradients()
        return g_fwd, g_bwd, grads
    except Exception as e:                                   # pragma: no cover - depends on driver / torch build
        import warnings
        warnings.warn(f"CUDA graph capture of the quantization step failed, running eagerly: {e}")
        try:
            torch.cuda.synchronize(device)
        except Exception:
            pass
        return False



def optimize_quantization_points(modelToQuantize, train_loader, test_loader, initial_learning_rate=1e-5,
                                 initial_momentum=0.9, epochs_to_train=30, print_every=500, use_nesterov=True,
                                 learning_rate_style="generic", numPointsPerTensor=16, assignBitsAutomatically=False,
                                 bucket_size=None, use_distillation_loss=True, initialize_method="quantiles",
                                 quantize_first_and_last_layer=True, *, max_steps=None, verbose=True, evaluate=True,
                                 step_hook=None, use_cuda_graphs=True, cuda_graph_step=False, use_plan=True):
    """Learn the quantization points of every tensor by SGD on the loss of the
    quantized network, the unquantized network acting as teacher (reference
    :395-592).  Returns ``(quantizedModel.state_dict(), pointsPerTensor, informationDict)``."""
    numTensorsNetwork = sum(1 for _ in modelToQuantize.parameters())
    initialize_method = initialize_method.lower()
    if initialize_method not in ("quantiles", "uniform"):
        raise ValueError("The initialization method must be either quantiles or uniform")
    if isinstance(numPointsPerTensor, int):
        numPointsPerTensor = [numPointsPerTensor] * numTensorsNetwork
    if len(numPointsPerTensor) != numTensorsNetwork:
        raise ValueError("numPointsPerTensor must be equal to the number of tensor in the network")
    if quantize_first_and_last_layer is False:
        numPointsPerTensor = numPointsPerTensor[1:-1]
    device = cnn_hf._device_of(modelToQuantize)
    scalingFunction = quantization.ScalingFunction("linear", False, False, bucket_size, False)     # :420

    if assignBitsAutomatically:                                             # :424-448
        num_to_estimate_grad = 5
        modelToQuantize.zero_grad()
        for idx_minibatch, batch in enumerate(train_loader, start=1):
            cnn_hf.forward_and_backward(modelToQuantize, batch, idx_batch=idx_minibatch, epoch=0,
                                        use_distillation_loss=False, return_tensor=True)
            if idx_minibatch >= num_to_estimate_grad:
                break
        # ||grad / num||_2 of every selected tensor: one multi-tensor launch, one device->host copy
        sel_grads = [p.grad for p in _selected_parameters(modelToQuantize, quantize_first_and_last_layer)]
        norms = (quantization.help_functions.gradient_norms(sel_grads) / num_to_estimate_grad).tolist()
        modelToQuantize.zero_grad()
        numPointsPerTensor = quantization.help_functions.assign_bits_automatically(norms, numPointsPerTensor,
                                                                                   input_is_point=True)

    selected = _selected_parameters(modelToQuantize, quantize_first_and_last_layer)
    pointsPerTensor = []
    # every tensor's points are a row of ONE (tensors x width) table padded with +inf, so that the
    # per-step re-sort of all lists (:550-551) is one torch.sort instead of one per tensor
    width = max(int(num) for num in numPointsPerTensor)
    points_table = torch.full((len(selected), width), float("inf"), dtype=torch.float32, device=device)
    for row, (p, num) in enumerate(zip(selected, numPointsPerTensor)):       # :451-482
        if initialize_method == "quantiles":
            init = quantization.help_functions.initialize_quantization_points(p.data, scalingFunction, num)
        else:
            init = torch.tensor([x / (num - 1) for x in range(num)], dtype=torch.float32, device=device)
        points_table[row, :num] = init.to(device)
        init = points_table[row, :num].detach().requires_grad_(True)         # leaf tensor sharing the table's storage
        init.grad = torch.zeros_like(init)
        pointsPerTensor.append(init)

    options = {"momentum": initial_momentum, "nesterov": use_nesterov} if initial_momentum != 0 else {}
    optimizer = optim.SGD(pointsPerTensor, lr=initial_learning_rate, **options)
    lr_scheduler = cnn_hf.LearningRateScheduler(initial_learning_rate, learning_rate_style)
    start_time = time.time()
    pred_accuracy_epochs, losses_epochs = [], []
    last_loss_saved = float("inf")
    batches_per_epoch = len(train_loader)
    if print_every > batches_per_epoch:
        print_every = max(batches_per_epoch // 2, 1)

    modelToQuantize.eval()
    quantizedModel = copy.deepcopy(modelToQuantize)                           # :497-498
    q_selected = _selected_parameters(quantizedModel, quantize_first_and_last_layer)
    quantizationFunctions = [quantization.nonUniformQuantization_variable(
        max_element=False, subtract_mean=False, modify_in_place=False, bucket_size=bucket_size,
        pre_process_tensors=True, tensor=p.data) for p in q_selected]         # :501-511

    # Multi-tensor plan: ONE launch quantizes every tensor with its current points, TWO produce every
    # centroid gradient (3 launches per step instead of 3 per tensor).  More than 32 points per tensor
    # or rows longer than 1024 elements: per-tensor ops (NotImplementedError from the plan).
    plan = None
    if use_plan and device.type == "cuda":
        try:
            plan = CentroidPlan([fun._tensor for fun in quantizationFunctions], [p.data for p in q_selected],
                                [pts.data for pts in pointsPerTensor], bucket_size)
        except NotImplementedError:
            plan = None

    def quantize_all():                                                       # :525-532
        if plan is not None:
            plan.forward_()
            return
        for fun, p_q, pts in zip(quantizationFunctions, q_selected, pointsPerTensor):
            fun.forward(None, pts.data, out=p_q.data)

    def point_gradients():                                                    # :539-545
        if plan is not None:
            return plan.backward_([p_q.grad.data if p_q.grad.is_contiguous() else p_q.grad.data.contiguous() for p_q in q_selected])
        return [fun.backward(p_q.grad.data)[1] for fun, p_q in zip(quantizationFunctions, q_selected)]

    # Without the plan the per-step quantization work is 3 small launches per tensor (22-60 tensors),
    # launch bound: after two eager steps the forward and the backward sequences are each captured
    # into a CUDA graph and replayed with one call per step.
    graphs = None
    graph_after = 2 if (use_cuda_graphs and plan is None and device.type == "cuda" and not cuda_graph_step) else None

    def one_step(data, idx_minibatch=1, epoch=0):
        """One step of the reference loop (:518-551)."""
        nonlocal graphs, graph_after
        quantizedModel.zero_grad(set_to_none=False)
        optimizer.zero_grad(set_to_none=False)
        if graphs is None and graph_after is not None and total_steps >= graph_after:
            graphs = _capture_point_graphs(quantize_all, point_gradients, device)
            if graphs is False:
                graph_after, graphs = None, None                          # capture unavailable: stay eager
        if graphs:
            graphs[0].replay()
        else:
            quantize_all()
        loss = cnn_hf.forward_and_backward(quantizedModel, data, idx_minibatch, epoch,
                                           use_distillation_loss=use_distillation_loss, teacher_model=modelToQuantize,
                                           return_tensor=True)
        if graphs:
            graphs[1].replay()
            grads = graphs[2]
        else:
            grads = point_gradients()
        for pts, gp in zip(pointsPerTensor, grads):
            pts.grad = gp
        optimizer.step()
        points_table.copy_(torch.sort(points_table, dim=1)[0])            # :550-551, every list at once, in place
        return loss, 0, 0

    # whole-step capture (opt-in), same mechanism as train_model(cuda_graph_step=True)
    whole_ok = bool(cuda_graph_step and device.type == "cuda")
    side_stream = torch.cuda.Stream(device) if whole_ok else None
    graphed = None

    total_steps, epoch, stop = 0, 0, False
    for epoch in range(epochs_to_train):
        quantizedModel.train()
        running = torch.zeros((), device=device)
        for idx_minibatch, data in enumerate(train_loader, start=1):
            if whole_ok and graphed is None and total_steps >= 3:
                graphed = _GraphedStep(one_step, data, device, side_stream, optimizer)
                if not graphed.ok:
                    whole_ok, graphed = False, None
            if graphed is not None and graphed.matches(data):
                loss = graphed.run(data)[0]
            elif whole_ok:
                with torch.cuda.stream(side_stream):
                    loss = one_step(data, idx_minibatch, epoch)[0]
                torch.cuda.current_stream(device).wait_stream(side_stream)
            else:
                loss = one_step(data, idx_minibatch, epoch)[0]
            running += loss
            total_steps += 1
            if step_hook is not None:
                step_hook(total_steps, loss)
            if idx_minibatch % print_every == 0:
                last_loss_saved = float(running.item()) / print_every
                running.zero_()
                if verbose:
                    print("Time Elapsed: {:.1f}s, [Epoch: {}, Minibatch: {}], loss: {:3f}".format(
                        time.time() - start_time, epoch + 1, idx_minibatch, last_loss_saved))
            if max_steps is not None and total_steps >= max_steps:
                stop = True
                break
        losses_epochs.append(last_loss_saved)
        if evaluate:
            pred_accuracy_epochs.append(cnn_hf.evaluateModel(quantizedModel, test_loader, fastEvaluation=False))
            if verbose:
                print(" === Epoch: {} - prediction accuracy {:2f}% === ".format(epoch + 1, pred_accuracy_epochs[-1] * 100))
        if stop:
            break
        error = 1 - pred_accuracy_epochs[-1] if pred_accuracy_epochs else 1.0
        new_learning_rate, stop_training = lr_scheduler.update_learning_rate(epoch, error)
        if stop_training is True:
            break
        for group in optimizer.param_groups:
            if group["lr"] != new_learning_rate:
                graphed = None                                             # the captured update holds the old rate
            group["lr"] = new_learning_rate
    informationDict = {"predictionAccuracy": pred_accuracy_epochs, "numEpochsTrained": epoch + 1,
                       "lossSaved": losses_epochs, "numStepsTrained": total_steps,
                       "cuda_graph_step": graphed is not None, "cuda_graph_quantization": bool(graphs),
                       "multi_tensor_plan": plan is not None}
    # the state dict also carries the batch-norm running statistics of the quantized model (:579-592)
    return quantizedModel.state_dict(), pointsPerTensor, informationDict
