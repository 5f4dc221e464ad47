"""GT jitter + repeat of SetCriterion.forward in training mode (conditional_detr.py:410-431) on the device
(csrc/targets.cu).  The random stream cannot be torch's; parity = the reference's acceptance rule, ordering rule and distribution:
  * every emitted row is either the original box or a candidate box * s with s in [1-j, 1+j)^4 and IoU(candidate, box) > 0.7;
  * per GT box the jittered rows come first, the original fills the rest, the LAST row is always the original (at most r-1 jittered);
  * labels / scores repeated r times, offsets scaled by r;
  * the per-coordinate scale statistics of accepted candidates match the reference's host loop (same rule, torch RNG);
  * a captured graph draws NEW candidates on every replay."""
import copy

import pytest
import torch

pytestmark = pytest.mark.gpu


def _targets(seed, B=6, with_scores=False):
    g = torch.Generator().manual_seed(seed)
    out = []
    for b in range(B):
        n = int(torch.randint(0 if b == 2 else 1, 9, (1,), generator=g)) if b != 2 else 0       # image 2 has no GT
        c = torch.rand(n, 2, generator=g) * 0.6 + 0.2
        wh = torch.rand(n, 2, generator=g) * 0.3 + 0.05
        t = {"labels": torch.randint(1, 81, (n,), generator=g), "boxes": torch.cat([c, wh], 1)}
        if with_scores:
            t["scores"] = torch.rand(n, generator=g)
        out.append(t)
    return out


def _iou(a, b):
    from oracle import spe_oracle as O
    return torch.diag(O.box_iou(O.box_cxcywh_to_xyxy(a), O.box_cxcywh_to_xyxy(b))[0])


@pytest.mark.parametrize("ratio,with_scores", [(5, False), (5, True), (1, False), (3, True)])
def test_device_jitter_repeat_rules(ratio, with_scores):
    from spe_b200 import criterion_ops as CO
    dev = torch.device("cuda")
    tg = _targets(3, with_scores=with_scores)
    T = CO.pack_targets(tg, dev)
    rng = CO.JitterRng(dev, seed=1234)
    E = CO.jitter_repeat(T, ratio, 0.1, rng)
    torch.cuda.synchronize()
    sizes = [len(t["labels"]) for t in tg]
    assert E.sizes == [n * ratio for n in sizes] and E.total == sum(sizes) * ratio
    assert E.offsets.cpu().tolist() == [o * ratio for o in T.offsets.cpu().tolist()]
    assert E.counts.cpu().tolist() == [n * ratio for n in sizes]
    n = T.total
    boxes = E.boxes[:n * ratio].cpu().view(n, ratio, 4)
    orig = T.boxes[:n].cpu()
    assert torch.equal(E.labels[:n * ratio].cpu().view(n, ratio), T.labels[:n].cpu().view(n, 1).expand(n, ratio))
    if with_scores:
        assert torch.equal(E.scores[:n * ratio].cpu().view(n, ratio), T.scores[:n].cpu().view(n, 1).expand(n, ratio))
    assert torch.equal(boxes[:, -1], orig)                                    # at most r-1 jittered rows: the last is the box itself
    for j in range(n):
        is_orig = (boxes[j] == orig[j]).all(1)
        k = int((~is_orig).sum())
        assert not is_orig[:k].any() and is_orig[k:].all()                    # jittered rows first, then the original
        if k:
            s = boxes[j, :k] / orig[j]
            assert float(s.min()) >= 0.9 - 1e-6 and float(s.max()) < 1.1 + 1e-6
            assert float(_iou(boxes[j, :k], orig[j].expand(k, 4)).min()) > 0.7 - 1e-5
    if ratio > 1:
        assert int((boxes[:, 0] != orig).any(1).sum()) >= 0.8 * n             # 1000 tries: (almost) every box finds an accepted copy
    # second call: the launch counter advanced -> different candidates, same rules
    E2 = CO.jitter_repeat(T, ratio, 0.1, rng)
    if ratio > 1:
        assert not torch.equal(E2.boxes[:n * ratio].cpu(), E.boxes[:n * ratio].cpu())
    assert int(rng.state[1]) == 2


def test_device_jitter_distribution_matches_reference_rule():
    """accepted-scale statistics: device kernel vs the host loop that replays the reference's draws (SetCriterion._jitter_repeat)."""
    from oracle import spe_oracle as O
    from spe_b200 import criterion_ops as CO, factory
    dev = torch.device("cuda")
    crit = factory.build_criterion(O.tiny_config(), device=dev, match_ratio=5).train()
    g = torch.Generator().manual_seed(9)
    n = 400
    tg = [{"labels": torch.ones(n, dtype=torch.int64), "boxes": torch.cat([torch.rand(n, 2, generator=g) * 0.6 + 0.2, torch.rand(n, 2, generator=g) * 0.3 + 0.05], 1)}]
    torch.manual_seed(0)
    host = crit._jitter_repeat(copy.deepcopy(tg))[0]["boxes"].view(n, 5, 4)
    E = CO.jitter_repeat(CO.pack_targets(tg, dev), 5, 0.1, CO.JitterRng(dev, seed=7))
    mine = E.boxes[:n * 5].cpu().view(n, 5, 4)
    sh = (host[:, :4] / tg[0]["boxes"].view(n, 1, 4)).reshape(-1, 4)
    sm = (mine[:, :4] / tg[0]["boxes"].view(n, 1, 4)).reshape(-1, 4)
    # 1600 accepted candidates each: means ~1, std ~ 0.1/sqrt(3) shrunk by the IoU > 0.7 acceptance -- both sides agree to sampling noise
    assert float((sh.mean(0) - sm.mean(0)).abs().max()) < 6e-3, (sh.mean(0), sm.mean(0))
    assert float((sh.std(0) - sm.std(0)).abs().max()) < 4e-3, (sh.std(0), sm.std(0))


def test_train_mode_criterion_uses_device_jitter_and_graph_replays_redraw():
    from oracle import spe_oracle as O
    from spe_b200 import factory
    from spe_b200.engine import TrainStep
    dev = torch.device("cuda")
    cfg = O.tiny_config()
    model = factory.build_detector(cfg, dev).train()
    model.load_state_dict(O.make_params(cfg, 5))
    crit = factory.build_criterion(cfg, device=dev, match_ratio=5).train()
    crit_r = factory.build_criterion(cfg, refine=True, device=dev, match_ratio=5).train()
    images, targets = O.make_inputs(cfg, 2, 48, 64, seed=5, max_gt=3)
    tgd = [{k: v.to(dev) for k, v in t.items()} for t in targets]
    trd = [dict(t, scores=torch.full((len(t["labels"]),), 0.6, device=dev)) for t in tgd]
    for graph in (False, True):
        step = TrainStep(model, crit, crit_r, graph=graph, max_gt=20)
        losses = []
        for _ in range(3):
            loss, ld, ld2 = step(images.to(dev), tgd, trd)
            losses.append(float(loss))
            assert torch.isfinite(loss)
        # jitter changes the GT every step -> the loss moves (same parameters, same images), in both modes
        assert len({round(l, 6) for l in losses}) == 3, (graph, losses)
        if graph:
            st = next(iter(step._g.values()))
            n = sum(len(t["labels"]) for t in targets)
            assert st["T"].total == 5 * n and st["Traw"].total == n
            lab = st["T"].labels[:5 * n].cpu().view(n, 5)
            assert torch.equal(lab, torch.cat([t["labels"] for t in targets]).to(torch.int32).view(n, 1).expand(n, 5))
    crit.device_jitter = False                                                # the reference's host loop is still available
    torch.manual_seed(1)
    ld = crit(model(images.to(dev)), tgd)
    assert torch.isfinite(ld["loss_ce"])
