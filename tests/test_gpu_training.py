"""GPU tests of the callers of the hot path: the quantized-distillation loop and the
differentiable-quantization loop (reference: cnn_models/conv_forward_model.py:165-592)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def env():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import quantized_distillation_b200.quantization as Q
    from quantized_distillation_b200.cnn_models import conv_forward_model as cfm
    from quantized_distillation_b200.cnn_models import help_fun as hf
    return Q, cfm, hf


def make_student(cfm):
    spec = dict(cfm.smallerModelSpec)
    spec["spec_dropout_rates"] = []
    return cfm.ConvolForwardNet(**spec, useBatchNorm=True, useAffineTransformInBatchNorm=True).cuda()


def distinct_per_bucket(t, bucket):
    flat = t.detach().view(-1)
    n = flat.numel()
    worst = 0
    for start in range(0, n, bucket):
        worst = max(worst, int(torch.unique(flat[start:start + bucket]).numel()))
    return worst


def test_weight_quantizer_matches_reference_choreography(env):
    """quantize -> restore -> gradient fix-up of the plan == the reference's per-tensor loop
    (conv_forward_model.py:236-266, 286, 302) done with the per-tensor API."""
    Q, cfm, hf = env
    torch.manual_seed(0)
    for first_last in (True, False):
        model = make_student(cfm)
        original = [p.detach().clone() for p in model.parameters()]
        wq = cfm.WeightQuantizer(model, numBits=4, bucket_size=256, backprop_quantization_style="complicated",
                                 quantize_first_and_last_layer=first_last)
        wq.quantize_weights_model()
        n_params = len(original)
        for i, (p, o) in enumerate(zip(model.parameters(), original)):
            if first_last is False and i in (0, n_params - 1):
                assert torch.equal(p, o)                              # skipped tensors untouched
            else:
                assert torch.equal(p, Q.uniformQuantization(o, 16, bucket_size=256)[0])
        wq.restore_weights_model()
        for p, o in zip(model.parameters(), original):
            assert torch.equal(p, o)
        for p in model.parameters():
            p.grad = torch.randn_like(p)
        grads = [p.grad.clone() for p in model.parameters()]
        wq.backward_quant_weights_model()
        for i, (p, o, g) in enumerate(zip(model.parameters(), original, grads)):
            if first_last is False and i in (0, n_params - 1):
                assert torch.equal(p.grad, g)
            else:
                f = Q.uniformQuantization_variable(16, bucket_size=256)
                f.forward(o)
                assert torch.equal(p.grad, f.backward(g))


@pytest.mark.parametrize("style", ["none", "truncated", "complicated"])
def test_quantized_distillation_runs_and_returns_quantized_weights(env, style):
    Q, cfm, hf = env
    torch.manual_seed(0)
    student = make_student(cfm)
    teacher = make_student(cfm).eval()
    data = hf.synthetic_cifar_loader(6, 25, seed=1)
    before = [p.detach().clone() for p in student.parameters()]
    losses = []
    model, info = cfm.train_model_quantized(student, data, data, numBits=4, bucket_size=256, use_distillation_loss=True,
                                            teacher_model=teacher, epochs_to_train=1, print_every=2, verbose=False,
                                            backprop_quantization_style=style, quantize_first_and_last_layer=False,
                                            step_hook=lambda i, l: losses.append(l))
    assert info["numStepsTrained"] == 6 and info["errorFlag"] is False
    assert all(torch.isfinite(l) for l in losses)
    params = list(model.parameters())
    assert any(not torch.equal(a, b) for a, b in zip(params, before))       # it trained
    for i, p in enumerate(params[1:-1], start=1):                            # final weights are quantized (:385)
        assert distinct_per_bucket(p, 256) <= 16, i
    assert distinct_per_bucket(params[0], 256) > 16                          # first tensor was skipped


def test_one_step_equals_manual_reference_step(env):
    """A full train_model step equals: quantize every tensor with the per-tensor API, forward/backward,
    restore, SGD -- the reference's step (:280-317) -- up to cuDNN's run-to-run float noise."""
    Q, cfm, hf = env
    torch.backends.cudnn.deterministic = True
    torch.manual_seed(3)
    a = make_student(cfm)
    b = make_student(cfm)
    b.load_state_dict(a.state_dict())
    teacher = make_student(cfm).eval()
    data = hf.synthetic_cifar_loader(1, 25, seed=5)
    cfm.train_model_quantized(a, data, data, numBits=4, bucket_size=256, use_distillation_loss=True, teacher_model=teacher,
                              epochs_to_train=1, print_every=1, verbose=False, evaluate=False, max_steps=1)
    # manual step on b
    opt = torch.optim.SGD(b.parameters(), lr=0.001, nesterov=True, momentum=0.9, weight_decay=0.00022)
    b.train()
    saved = [p.detach().clone() for p in b.parameters()]
    for p in b.parameters():
        p.data = Q.uniformQuantization(p.data, 16, bucket_size=256)[0]
    b.zero_grad()
    hf.forward_and_backward(b, data[0], 1, 0, use_distillation_loss=True, teacher_model=teacher)
    for p, s in zip(b.parameters(), saved):
        p.data = s
    opt.step()
    for p in b.parameters():                                                 # train_model returns quantized weights
        p.data = Q.uniformQuantization(p.data, 16, bucket_size=256)[0]
    for pa, pb in zip(a.parameters(), b.parameters()):
        assert torch.allclose(pa.detach(), pb.detach(), rtol=0, atol=float(pb.detach().abs().max()) * 0.08 + 1e-6)
        assert (pa != pb).float().mean() < 0.02                               # at most a few level flips from float noise
    torch.backends.cudnn.deterministic = False


def test_differentiable_quantization_loop(env):
    Q, cfm, hf = env
    torch.manual_seed(0)
    model = make_student(cfm)
    data = hf.synthetic_cifar_loader(6, 25, seed=2)
    state, points, info = cfm.optimize_quantization_points(
        model, data, data, initial_learning_rate=1e-3, epochs_to_train=1, print_every=2, numPointsPerTensor=4,
        bucket_size=256, use_distillation_loss=True, initialize_method="quantiles", verbose=False)
    assert info["numStepsTrained"] == 6 and len(points) == 22
    sf = Q.ScalingFunction("linear", False, False, 256, False)
    for p, (name, w) in zip(points, [(k, v) for k, v in state.items() if k in dict(model.named_parameters())]):
        assert p.numel() == 4 and bool((p[1:] >= p[:-1]).all())                # kept sorted (:550-551)
    # quantized model weights only take centroid values (in scaled space) of their tensor
    params = dict(model.named_parameters())
    checked = 0
    for (name, p0), pts in zip(params.items(), points):
        wq = state[name]
        f = Q.nonUniformQuantization_variable(bucket_size=256, pre_process_tensors=True, tensor=p0.data)
        # the loop wrote forward(points BEFORE the last update); check the invariant instead: <= 4 values per bucket
        assert distinct_per_bucket(wq, 256) <= 4, name
        checked += 1
    assert checked == 22
    # uniform initialisation + automatic bit assignment path
    state, points, info = cfm.optimize_quantization_points(
        model, data, data, initial_learning_rate=1e-3, epochs_to_train=1, print_every=2, numPointsPerTensor=4,
        bucket_size=256, use_distillation_loss=True, initialize_method="uniform", assignBitsAutomatically=True,
        quantize_first_and_last_layer=False, verbose=False, max_steps=2, evaluate=False)
    assert len(points) == 20 and sum(p.numel() for p in points) == 80
    with pytest.raises(ValueError):
        cfm.optimize_quantization_points(model, data, data, initialize_method="kmeans")


def test_wide_resnet_quantized_step(env):
    Q, cfm, hf = env
    from quantized_distillation_b200.cnn_models.wide_resnet import Wide_ResNet
    torch.manual_seed(0)
    student = Wide_ResNet(depth=10, widen_factor=2, dropout_rate=0.3, num_classes=10).cuda()
    teacher = Wide_ResNet(depth=10, widen_factor=2, dropout_rate=0.0, num_classes=10).cuda().eval()
    data = hf.synthetic_cifar_loader(3, 16, seed=3)
    model, info = cfm.train_model_quantized(student, data, data, numBits=2, bucket_size=256, use_distillation_loss=True,
                                            teacher_model=teacher, initial_learning_rate=0.1, weight_decayL2=5e-4,
                                            learning_rate_style="cifar100", epochs_to_train=1, print_every=1, verbose=False,
                                            quantize_first_and_last_layer=False)
    assert info["numStepsTrained"] == 3
    params = list(model.parameters())
    assert distinct_per_bucket(params[5], 256) <= 4


def test_diffquant_cuda_graph_path_matches_eager(env):
    """The CUDA-graph replay of the per-step quantization launches gives the same centroids as eager launches."""
    Q, cfm, hf = env
    torch.backends.cudnn.deterministic = True
    results = []
    for use_graphs, whole in ((False, False), (True, False), (False, True)):
        torch.manual_seed(11)
        model = make_student(cfm)
        data = hf.synthetic_cifar_loader(8, 25, seed=9)
        state, points, info = cfm.optimize_quantization_points(
            model, data, data, initial_learning_rate=1e-3, epochs_to_train=1, print_every=4, numPointsPerTensor=4,
            bucket_size=256, use_distillation_loss=True, initialize_method="quantiles", verbose=False, evaluate=False,
            use_cuda_graphs=use_graphs, cuda_graph_step=whole)
        assert info["numStepsTrained"] == 8
        results.append([p.detach().clone() for p in points])
    torch.backends.cudnn.deterministic = False
    for a, b, c in zip(*results):
        assert torch.allclose(a, b, rtol=1e-4, atol=1e-6), (a, b)
        assert torch.allclose(a, c, rtol=1e-4, atol=1e-6), (a, c)
    assert any(not torch.equal(a, b0) for a, b0 in zip(results[0], [torch.zeros_like(x) for x in results[0]]))


@pytest.mark.parametrize("style", ["none", "complicated"])
def test_whole_step_cuda_graph_matches_eager(env, style):
    """cuda_graph_step=True replays exactly the eager step: same weights after 8 steps (up to cuDNN float noise)."""
    Q, cfm, hf = env
    torch.backends.cudnn.deterministic = True
    finals = []
    for graph in (False, True):
        torch.manual_seed(21)
        student = make_student(cfm)
        teacher = make_student(cfm).eval()
        data = hf.synthetic_cifar_loader(8, 25, seed=4)
        model, info = cfm.train_model_quantized(student, data, data, numBits=4, bucket_size=256, use_distillation_loss=True,
                                                teacher_model=teacher, epochs_to_train=1, print_every=4, verbose=False,
                                                evaluate=False, backprop_quantization_style=style, cuda_graph_step=graph)
        assert info["numStepsTrained"] == 8
        finals.append([p.detach().clone() for p in model.parameters()])
    torch.backends.cudnn.deterministic = False
    # Replay runs the very kernels the eager step launches, so the trajectories can only part where cuDNN
    # picks another algorithm under capture; after 8 steps at lr 1e-3 that moves a weight by ~1e-7, which
    # flips its 4-bit level only if it sat on a rounding boundary: at most a handful per tensor.
    for a, b in zip(*finals):
        frac = float((a != b).float().mean())
        step = float((a.max() - a.min()) / 15) if a.numel() > 1 else 1.0
        assert frac <= 0.002, f"{frac:.5f} of the weights differ between graph and eager training"
        assert float((a - b).abs().max()) <= 1.01 * step + 1e-6, "a weight moved by more than one quantization level"


@pytest.mark.parametrize("graph", [False, True])
def test_mix_with_differentiable_quantization_two_epochs(env, graph):
    """train_model(mix_with_differentiable_quantization=True): after every epoch but the last the
    quantization points are optimised for one epoch and the result is loaded back (reference
    cnn_models/conv_forward_model.py:342-353).  Round 1 crashed on the second epoch (the step-state
    dict was overwritten by the returned state dict)."""
    Q, cfm, hf = env
    torch.manual_seed(5)
    student = make_student(cfm)
    teacher = make_student(cfm).eval()
    data = hf.synthetic_cifar_loader(4, 25, seed=9)
    before = [p.detach().clone() for p in student.parameters()]
    model, info = cfm.train_model_quantized(student, data, data, numBits=2, bucket_size=256, use_distillation_loss=True,
                                            teacher_model=teacher, epochs_to_train=2, print_every=2, verbose=False,
                                            evaluate=False, mix_with_differentiable_quantization=True, cuda_graph_step=graph)
    assert info["errorFlag"] is False and info["numStepsTrained"] == 8
    assert info["numEpochsTrained"] == 4                              # doubled like the reference (:376-377)
    assert len(info["lossSaved"]) == 3                                # epoch 1, the differentiable epoch, epoch 2
    params = list(model.parameters())
    assert any(not torch.equal(a, b) for a, b in zip(before, params))
    for p in params:                                                   # returned weights are 2-bit quantized per bucket
        assert distinct_per_bucket(p, 256) <= 4
        assert torch.isfinite(p).all()


def test_diffquant_multi_tensor_plan_matches_per_tensor_loop(env):
    """optimize_quantization_points with the CentroidPlan (3 launches per step for the whole model)
    against the same loop on the per-tensor ops (3 launches per tensor): same quantized weights bit
    for bit at every step, points equal up to the summation order of the centroid gradients."""
    Q, cfm, hf = env
    torch.backends.cudnn.deterministic = True
    outs = []
    for use_plan in (True, False):
        torch.manual_seed(3)
        model = make_student(cfm)
        data = hf.synthetic_cifar_loader(5, 25, seed=6)
        state, points, info = cfm.optimize_quantization_points(
            model, data, data, initial_learning_rate=1e-4, epochs_to_train=1, print_every=5, numPointsPerTensor=4,
            bucket_size=256, use_distillation_loss=True, initialize_method="quantiles", verbose=False, evaluate=False,
            use_cuda_graphs=False, use_plan=use_plan)
        assert info["multi_tensor_plan"] is use_plan and info["numStepsTrained"] == 5
        outs.append(([p.detach().clone() for p in points], state))
    torch.backends.cudnn.deterministic = False
    for a, b in zip(outs[0][0], outs[1][0]):
        assert torch.allclose(a, b, rtol=1e-5, atol=1e-7), (a, b)
    agree = [float((outs[0][1][k] == outs[1][1][k]).float().mean()) for k in outs[0][1] if outs[0][1][k].dtype == torch.float32]
    assert min(agree) > 0.999, min(agree)                     # a point moving by an ulp can flip a boundary element


@pytest.mark.parametrize("style", ["none", "truncated", "complicated"])
def test_fused_sgd_step_is_torch_sgd_then_quantize_bit_for_bit(env, style):
    """qd_plan_sgd_step (f1, second half): gradient fix-up + torch.optim.SGD(momentum, Nesterov, weight decay)
    + re-quantization in one pass.  Against the unfused chain -- the plan's fix-up launch, torch.optim.SGD
    itself on the full-precision weights, uniformQuantization per tensor -- master, momentum buffer and
    quantized weights must agree BIT FOR BIT over several steps (this pins the FMA policy of the kernel
    to the one torch's CUDA kernels compile to)."""
    Q, cfm, hf = env
    from quantized_distillation_b200.plan import QuantizationPlan
    gen = torch.Generator(device="cuda").manual_seed(77)
    sizes = [5000, 10, 5625, 75, 93750, 50, 800000, 500, 257, 255, 1]
    scale = 0.6 if style == "truncated" else 0.05
    live = [torch.randn(n, generator=gen, device="cuda") * scale for n in sizes]
    ref_w = [torch.nn.Parameter(t.clone()) for t in live]
    lr, mu, wd = 1e-2, 0.9, 2.2e-4
    opt = torch.optim.SGD(ref_w, lr=lr, momentum=mu, nesterov=True, weight_decay=wd)
    plan = QuantizationPlan(live, 16, 256)
    ref_plan = QuantizationPlan([p.data for p in ref_w], 16, 256)       # only its fix-up launch is used, on the fp32 weights
    if style == "truncated":
        for t in live:
            t.clamp_(-1, 1)
    plan.save_and_quantize_()
    for step in range(4):
        grads = [torch.randn(n, generator=gen, device="cuda") for n in sizes]
        # ---- unfused reference chain ----
        if style == "truncated":
            for p in ref_w:
                p.data.clamp_(-1, 1)                                    # quantize-time clamp (:240-241)
        rg = [g.clone() for g in grads]
        ref_plan.backward_(rg, style)                                   # fix-up at the full-precision weights (:315)
        for p, g in zip(ref_w, rg):
            p.grad = g
        opt.step()
        if style == "truncated":
            for p in ref_w:
                p.data.clamp_(-1, 1)
        # ---- fused ----
        plan.fused_step_([g.clone() for g in grads], style, lr, mu, wd, True)
        for i, p in enumerate(ref_w):
            assert torch.equal(plan._master[i].view(-1), p.data.view(-1)), (style, step, i, "master")
            buf = opt.state[p]["momentum_buffer"]
            assert bool((plan.momentum_buffers[i].view(-1) == buf.view(-1)).all()), (style, step, i, "momentum")
            q, _ = Q.uniformQuantization(p.data, 16, bucket_size=256)
            assert torch.equal(live[i].view(-1), q.view(-1)), (style, step, i, "quantized")
    with pytest.raises(NotImplementedError):
        big = QuantizationPlan([torch.randn(4096, device="cuda")], 16, 1024)
        big.fused_step_([torch.randn(4096, device="cuda")], "none", lr, mu, wd, True)


@pytest.mark.parametrize("style", ["none", "complicated"])
def test_fused_training_loop_equals_unfused(env, style):
    """train_model(fused_optimizer_step=True) walks through the same weights as the unfused loop."""
    Q, cfm, hf = env
    torch.backends.cudnn.deterministic = True
    finals = []
    for fused in (False, True):
        torch.manual_seed(21)
        student = make_student(cfm)
        teacher = make_student(cfm).eval()
        data = hf.synthetic_cifar_loader(6, 25, seed=4)
        model, info = cfm.train_model_quantized(student, data, data, numBits=4, bucket_size=256, use_distillation_loss=True,
                                                teacher_model=teacher, epochs_to_train=1, print_every=3, verbose=False,
                                                evaluate=False, backprop_quantization_style=style, fused_optimizer_step=fused,
                                                quantize_first_and_last_layer=False)
        assert info["fused_optimizer_step"] is fused and info["numStepsTrained"] == 6
        finals.append([p.detach().clone() for p in model.parameters()])
    torch.backends.cudnn.deterministic = False
    same = [float((a == b).float().mean()) for a, b in zip(*finals)]
    assert min(same) == 1.0, same


def test_absmax_quantized_training_extension(env):
    """quantizationFunctionToUse='uniformAbsMaxScaling' (reference :206-208, broken there): refused by default, runs as the
    opt-in unpinned extension; weights come back with at most 2^(numBits-1) magnitudes per bucket."""
    Q, cfm, hf = env
    from quantized_distillation_b200.quantization import quant_functions as QF
    torch.manual_seed(2)
    data = hf.synthetic_cifar_loader(3, 25, seed=1)
    with pytest.raises(NotImplementedError):
        cfm.train_model_quantized(make_student(cfm), data, data, numBits=3, bucket_size=256, evaluate=False, verbose=False,
                                  quantizationFunctionToUse="uniformAbsMaxScaling", epochs_to_train=1)
    QF.ALLOW_UNPINNED_SCALING = True
    try:
        model, info = cfm.train_model_quantized(make_student(cfm), data, data, numBits=3, bucket_size=256, evaluate=False,
                                                verbose=False, quantizationFunctionToUse="uniformAbsMaxScaling", epochs_to_train=1,
                                                backprop_quantization_style="truncated", print_every=1)
        assert info["numStepsTrained"] == 3
        for p in model.parameters():
            flat = p.detach().abs().view(-1)
            for start in range(0, flat.numel(), 256):
                assert torch.unique(flat[start:start + 256]).numel() <= 4          # s = 2^(3-1) magnitudes
        with pytest.raises(ValueError):
            cfm.train_model_quantized(make_student(cfm), data, data, numBits=3, bucket_size=256, evaluate=False, verbose=False,
                                      quantizationFunctionToUse="uniformAbsMaxScaling", backprop_quantization_style="complicated")
    finally:
        QF.ALLOW_UNPINNED_SCALING = False
