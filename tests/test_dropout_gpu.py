"""Dropout / DropPath / attention dropout (the reference's training configuration: scripts/run_coco17.py:30-32, main.py:73).
Attention-probability dropout is checked against fp32 autograd with the SAME keep-mask (ops._drop_mask patched to a recorded mask);
the module-level routes are checked for eval/p=0 equivalence with the fused routes and for a full train step with the script's
arguments through build_model(args)."""
import argparse
import json
import os

import pytest
import torch

pytestmark = pytest.mark.gpu


def rel(a, b):
    return float((a.float() - b.float()).norm() / (b.float().norm() + 1e-30))


def _fixed_mask(monkeypatch, K, p, seed):
    rec = {}

    def fake(shape, pp, device):
        g = torch.Generator(device="cpu").manual_seed(seed)
        rec["keep"] = (torch.rand(shape, generator=g) >= pp).to(device)
        return rec["keep"]

    monkeypatch.setattr(K, "_drop_mask", fake)
    return rec


@pytest.mark.parametrize("masked", [False, True])
def test_attention_probability_dropout_fixed_mask(monkeypatch, masked):
    from spe_b200 import ops as K
    g = torch.Generator().manual_seed(5)
    B, H, Lq, Lk, d, p = 2, 8, 70, 100, 48, 0.1
    dev = torch.device("cuda")
    mk = lambda L, dd: (torch.randn(B, L, H * dd, generator=g)).to(torch.bfloat16).to(dev).requires_grad_(True)
    q, k, v = mk(Lq, d), mk(Lk, d), mk(Lk, d)
    mask = None
    if masked:
        mask = torch.zeros(B, Lk, dtype=torch.uint8)
        mask[0, 90:] = 1
        mask = mask.to(dev)
    rec = _fixed_mask(monkeypatch, K, p, 77)
    out = K.attention(q, k, v, H, d ** -0.5, mask_u8=mask, drop_p=p)
    go = torch.randn(out.shape, generator=g).to(torch.bfloat16).to(dev)
    got = torch.autograd.grad((out.float() * go.float()).sum(), [q, k, v])
    keep = rec["keep"][..., :Lk]
    fr = [t.detach().float().requires_grad_(True) for t in (q, k, v)]
    hd = lambda t, L: t.reshape(B, L, H, d).transpose(1, 2)
    S = hd(fr[0], Lq) @ hd(fr[1], Lk).transpose(-1, -2) * d ** -0.5
    if masked:
        S = S.masked_fill(mask.bool()[:, None, None, :], float("-inf"))
    P = S.softmax(-1) * keep / (1 - p)
    outr = (P @ hd(fr[2], Lk)).transpose(1, 2).reshape(B, Lq, H * d)
    ref = torch.autograd.grad((outr * go.float()).sum(), fr)
    assert rel(out, outr) < 2e-2, rel(out, outr)
    for a, b in zip(got, ref):
        assert rel(a, b) < 3e-2, rel(a, b)
    # without dropout the same call takes the fused kernel and matches the keep-everything reference
    out0 = K.attention(q, k, v, H, d ** -0.5, mask_u8=mask)
    assert rel(out0, (S.softmax(-1) @ hd(fr[2], Lk)).transpose(1, 2).reshape(B, Lq, H * d)) < 2e-2


def test_talking_heads_attention_dropout_fixed_mask(monkeypatch):
    from spe_b200 import ops as K
    g = torch.Generator().manual_seed(6)
    B, H, N, dh, p = 2, 4, 150, 48, 0.05
    D = H * dh
    dev = torch.device("cuda")
    qkv = torch.randn(B, N, 3 * D, generator=g).to(torch.bfloat16).to(dev).requires_grad_(True)
    Wl = (torch.eye(H) + 0.3 * torch.randn(H, H, generator=g)).to(dev).requires_grad_(True)
    bl = (0.1 * torch.randn(H, generator=g)).to(dev).requires_grad_(True)
    Ww = (torch.eye(H) + 0.3 * torch.randn(H, H, generator=g)).to(dev).requires_grad_(True)
    bw = (0.05 * torch.randn(H, generator=g)).to(dev).requires_grad_(True)
    rec = _fixed_mask(monkeypatch, K, p, 78)
    out = K.talking_heads_attention(qkv, Wl, bl, Ww, bw, H, drop_p=p)
    go = torch.randn(out.shape, generator=g).to(torch.bfloat16).to(dev)
    got = torch.autograd.grad((out.float() * go.float()).sum(), [qkv, Wl, Ww, bw])
    keep = rec["keep"][..., :N]
    qr = qkv.detach().float().requires_grad_(True)
    ps = [t.detach().clone().requires_grad_(True) for t in (Wl, Ww, bw)]
    t = qr.view(B, N, 3, H, dh).permute(2, 0, 3, 1, 4)
    S = (t[0] * dh ** -0.5) @ t[1].transpose(-1, -2)
    L = torch.einsum("gh,bhij->bgij", ps[0], S) + bl.detach().view(1, H, 1, 1)
    A = (torch.einsum("gh,bhij->bgij", ps[1], L.softmax(-1)) + ps[2].view(1, H, 1, 1)) * keep / (1 - p)      # attn_drop after proj_w (cait.py:386-387)
    outr = (A @ t[2]).transpose(1, 2).reshape(B, N, D)
    ref = torch.autograd.grad((outr * go.float()).sum(), [qr] + ps)
    assert rel(out, outr) < 2e-2, rel(out, outr)
    for a, b, n in zip(got, ref, ["dqkv", "dWl", "dWw", "dbw"]):
        assert rel(a, b) < (6e-2 if n == "dbw" else 3e-2), (n, rel(a, b))      # dbw: the kernel's bf16 sum of masked dA


def test_drop_path_and_dropout_helpers():
    from spe_b200 import ops as K
    torch.manual_seed(0)
    x = torch.ones(4096, 3, 5, device="cuda")
    assert K.dropout(x, 0.3, False) is x and K.drop_path(x, 0.3, False) is x and K.dropout(x, 0.0, True) is x
    y = K.drop_path(x, 0.25, True)
    per = y.reshape(4096, -1)
    assert ((per == 0).all(1) | (per == 1 / 0.75).all(1)).all()               # whole samples are dropped or rescaled
    assert abs(float(y.mean()) - 1.0) < 0.05
    z = K.dropout(x, 0.1, True)
    assert abs(float(z.mean()) - 1.0) < 0.02 and abs(float((z == 0).float().mean()) - 0.1) < 0.01


def test_train_step_with_reference_script_args(golden_dir):
    """build_model(args) with scripts/run_coco17.py's arguments (XXS36 two-branch, drop rates 0.07 / 0.2 / 0.05, decoder dropout 0.1,
    hung_match_ratio 5): one training step in train() mode runs, is finite, reaches every parameter group; eval() is deterministic
    and equals the fused no-dropout route."""
    from spe_b200.models import build_model
    fix = json.load(open(os.path.join(golden_dir, "build_args.json")))
    d = dict(fix["cases"]["run_coco17"]["args"])
    d["device"] = "cuda"
    args = argparse.Namespace(**d)
    torch.manual_seed(0)
    model, criterion, criterion_refine, _, _ = build_model(args)
    dev = torch.device("cuda")
    model.to(dev)
    with torch.no_grad():                                      # LayerScale init 1e-5 would hide the backbone branches
        for n, p in model.named_parameters():
            if "gamma_" in n:
                p.fill_(0.1)
    g = torch.Generator().manual_seed(1)
    B = 2
    images = torch.randn(B, 3, 128, 160, generator=g).to(dev)
    targets = []
    for b in range(B):
        n = 2 + b
        boxes = torch.cat([torch.rand(n, 2, generator=g) * 0.5 + 0.25, torch.rand(n, 2, generator=g) * 0.3 + 0.05], 1)
        labels = torch.randint(1, 80, (n,), generator=g)
        il = torch.zeros(90)
        il[labels - 1] = 1
        targets.append({"labels": labels.to(dev), "boxes": boxes.to(dev), "img_label": il.to(dev),
                        "scores": (torch.rand(n, generator=g) * 0.8 + 0.1).to(dev)})
    model.train(); criterion.train(); criterion_refine.train()
    if getattr(args, "hungarian_multi", False):
        criterion.update_hung_match_ratio(args.hung_match_ratio)
        criterion_refine.update_hung_match_ratio(args.hung_match_ratio)
    out = model(images)
    ld = criterion(out[0], targets)
    ld2 = criterion_refine(out[1], targets)
    wd = criterion.weight_dict
    loss = sum(ld[k] * wd[k] for k in ld if k in wd) + sum(ld2[k] * wd[k] for k in ld2 if k in wd)
    assert torch.isfinite(loss)
    loss.backward()
    groups = {"backbone": 0.0, "blocks_token_only": 0.0, "transformer": 0.0, "blocks_det": 0.0}
    for n, p in model.named_parameters():
        if p.grad is None:
            continue
        assert torch.isfinite(p.grad).all(), n
        for k in groups:
            if k in n:
                groups[k] += float(p.grad.abs().sum())
    assert all(v > 0 for v in groups.values()), groups
    out_b = model(images)                                      # train mode: a second forward draws new masks
    assert not torch.equal(out_b[0]["pred_logits"], out[0]["pred_logits"])
    model.eval()
    with torch.no_grad():
        e1, e2 = model(images), model(images)
    # no masks in eval(): repeatable up to the order of the fp32 reduce-adds of the chunked kernels
    # (a bf16 rounding flip downstream of a last-bit difference moves single logits by up to ~1e-2)
    torch.testing.assert_close(e1[0]["pred_logits"], e2[0]["pred_logits"], rtol=1e-2, atol=3e-2)
    torch.testing.assert_close(e1[1]["pred_boxes"], e2[1]["pred_boxes"], rtol=1e-2, atol=5e-3)
    assert float((out[0]["pred_logits"] - e1[0]["pred_logits"]).abs().max()) > 1e-2           # and different from the dropped forward
