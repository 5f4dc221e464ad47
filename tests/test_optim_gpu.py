"""clip_grad_norm_ + AdamW with the reference's three parameter groups (main.py:177-190, engine.py:163-164) on the flat
buffers, against torch.optim.AdamW + torch.nn.utils.clip_grad_norm_ on a copy of the same model, several steps."""
import copy

import pytest
import torch

pytestmark = pytest.mark.gpu


def _groups(model, lr, lr_backbone, lr_cls_head):
    named = list(model.named_parameters())
    return [{"params": [p for n, p in named if "backbone" not in n and p.requires_grad]},
            {"params": [p for n, p in named if "backbone" in n and p.requires_grad and "blocks_token_only" not in n], "lr": lr_backbone},
            {"params": [p for n, p in named if "backbone" in n and p.requires_grad and "blocks_token_only" in n], "lr": lr_cls_head}]


@pytest.mark.parametrize("max_norm", [0.1, 0.0])
def test_flat_adamw_matches_torch(max_norm):
    from oracle import spe_oracle as O
    from spe_b200 import factory
    from spe_b200.dp import FlatGradBuffer
    from spe_b200.optim import FlatAdamW
    dev = torch.device("cuda")
    cfg = O.tiny_config()
    params = O.make_params(cfg, 3)
    model = factory.build_detector(cfg, dev)
    model.load_state_dict(params)
    ref = copy.deepcopy(model)
    lr, lrb, lrc, wd = 1e-3, 2e-4, 5e-4, 1e-2
    opt_ref = torch.optim.AdamW(_groups(ref, lr, lrb, lrc), lr=lr, weight_decay=wd)
    gbuf = FlatGradBuffer(model.parameters())
    opt = FlatAdamW(model, gbuf, lr=lr, lr_backbone=lrb, lr_cls_head=lrc, weight_decay=wd, clip_max_norm=max_norm, write_clipped_grads=True)
    # the three groups partition the buffer exactly as the reference's name filters do
    n_by_group = [0, 0, 0]
    prev = 0
    for end, g in opt.segments:
        n_by_group[g] += end - prev
        prev = end
    want = [sum(p.numel() for p in grp["params"]) for grp in _groups(model, lr, lrb, lrc)]
    assert all(a >= b and a - b < 8 * len(list(model.parameters())) for a, b in zip(n_by_group, want)), (n_by_group, want)
    g = torch.Generator().manual_seed(0)
    for it in range(4):
        gbuf.zero_()
        for p, q in zip(model.parameters(), ref.parameters()):
            gr = torch.randn(p.shape, generator=g).to(dev) * (10.0 if it == 1 else 0.01)       # step 1 is clipped hard, the others are not
            p.grad.copy_(gr)
            q.grad = gr.clone()
        if max_norm > 0:
            tn = torch.nn.utils.clip_grad_norm_(ref.parameters(), max_norm)
        opt_ref.step()
        opt.step()
        if max_norm > 0:
            assert abs(opt.total_norm() - float(tn)) <= 1e-4 * float(tn)
            for p, q in zip(model.parameters(), ref.parameters()):              # write_clipped_grads: .grad holds what clip_grad_norm_ leaves
                assert torch.allclose(p.grad, q.grad, rtol=1e-4, atol=1e-9)
        if it == 1:
            opt.step_lr(epoch=1, lr_drop=1)                                       # StepLR drop between steps 1 and 2
            for grp in opt_ref.param_groups:
                grp["lr"] *= 0.1
    worst = 0.0
    for (n, p), q in zip(model.named_parameters(), ref.parameters()):
        d = float((p.detach() - q.detach()).abs().max())
        s = float((q.detach() - params[n].to(dev)).abs().max())                   # size of the accumulated update
        worst = max(worst, d / (s + 1e-12))
        pmax = float(q.detach().abs().max())
        assert d <= 1e-4 * s + 5e-7 * pmax + 1e-9, (n, d, s, pmax)      # fp32 rounding: fma contraction / bias-correction powers differ by ulps
    print("flat AdamW vs torch.optim.AdamW: worst |dp| / |update| = %.2e" % worst)
    # the model still runs on the re-pointed parameters and sees the updated weights (bf16 shadows refreshed by the step)
    images, _ = O.make_inputs(cfg, 1, 48, 64, seed=1)
    out = model(images.to(dev))
    out_ref = ref(images.to(dev))
    assert torch.allclose(out[0]["pred_logits"], out_ref[0]["pred_logits"], rtol=2e-2, atol=2e-2)


def test_clip_grad_norm_flat():
    from spe_b200.dp import FlatGradBuffer
    from spe_b200.optim import clip_grad_norm_
    dev = torch.device("cuda")
    ps = [torch.nn.Parameter(torch.randn(s, device=dev)) for s in [(33, 7), (5,), (128, 64), (3, 3, 3)]]
    qs = [torch.nn.Parameter(p.detach().clone()) for p in ps]
    gbuf = FlatGradBuffer(ps)
    for p, q in zip(ps, qs):
        g = torch.randn_like(p)
        p.grad.copy_(g)
        q.grad = g.clone()
    tn_ref = torch.nn.utils.clip_grad_norm_(qs, 0.1)
    tn = clip_grad_norm_(gbuf, 0.1)
    assert abs(float(tn) - float(tn_ref)) <= 1e-5 * float(tn_ref)
    for p, q in zip(ps, qs):
        assert torch.allclose(p.grad, q.grad, rtol=1e-5, atol=1e-9)


@pytest.mark.parametrize("graph", [False, True])
def test_train_step_with_fused_optimizer_matches_reference_loop(graph):
    """engine.py:161-165 (zero_grad, backward, clip_grad_norm_, optimizer.step) as ONE TrainStep call (captured in the CUDA graph when
    graph=True).  After every call the gradients the step produced (still in the flat buffer) are handed to torch's
    clip_grad_norm_ + AdamW on a shadow copy of the parameters: the two parameter sets must stay equal step after step.  (Comparing
    two independently computed loss trajectories instead is meaningless: Adam's first updates are lr * sign(g), so the atomics-order
    noise of the backward flips noise-level gradient signs and the runs drift apart within three steps.)"""
    from oracle import spe_oracle as O
    from spe_b200 import factory
    from spe_b200.dp import FlatGradBuffer
    from spe_b200.engine import TrainStep
    from spe_b200.optim import FlatAdamW
    dev = torch.device("cuda")
    cfg = O.tiny_config()
    params = O.make_params(cfg, 11)
    images, targets = O.make_inputs(cfg, 2, 48, 64, seed=11, max_gt=3)
    tg = [{k: v.to(dev) for k, v in t.items()} for t in targets]
    lr, lrb, lrc, wd, clip = 2e-3, 1e-3, 1.5e-3, 1e-2, 0.1
    model = factory.build_detector(cfg, dev).train()
    model.load_state_dict(params)
    crit = factory.build_criterion(cfg, device=dev).eval()
    gbuf = FlatGradBuffer(model.parameters())
    opt = FlatAdamW(model, gbuf, lr=lr, lr_backbone=lrb, lr_cls_head=lrc, weight_decay=wd, clip_max_norm=clip)
    step = TrainStep(model, crit, None, grad_buffer=gbuf, graph=graph, max_gt=8, optimizer=opt)
    ref = copy.deepcopy(model)                      # parameter holder for torch's optimizer (never run forward)
    opt_ref = torch.optim.AdamW(_groups(ref, lr, lrb, lrc), lr=lr, weight_decay=wd)
    losses = []
    for it in range(4):
        losses.append(float(step(images.to(dev), tg)[0]))
        for p, q in zip(model.parameters(), ref.parameters()):
            q.grad = p.grad.detach().clone()        # the step's own gradients (the flat buffer keeps them unclipped)
        torch.nn.utils.clip_grad_norm_(ref.parameters(), clip)
        opt_ref.step()
        for (n, p), q in zip(model.named_parameters(), ref.parameters()):
            d = float((p.detach() - q.detach()).abs().max())
            assert d <= 1e-4 * lr * (it + 1) + 5e-7 * float(q.detach().abs().max()) + 1e-9, (graph, it, n, d)
    assert int(opt.state[3]) == 4                   # the capture warm-up passes did not step the optimizer
    assert losses[-1] < losses[0], losses           # and the model is learning
