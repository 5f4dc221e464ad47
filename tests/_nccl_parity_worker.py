"""Worker of tests/test_dist_nccl_gpu.py (launched with torch.distributed.run, one rank per GPU).
SURVEY section 4 item 4: batch 8 on one GPU vs 2 x 4 over NCCL -> same losses, same gradients after the all-reduce."""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import faulthandler
    faulthandler.dump_traceback_later(int(os.environ.get("SPE_TEST_DUMP_AFTER", "240")), exit=True)     # a hang prints where, then ends
    from oracle import spe_oracle as O
    from spe_b200 import factory
    from spe_b200.dp import FlatGradBuffer
    from spe_b200.engine import TrainStep
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", rank)))
    dev = torch.device("cuda")
    graph = os.environ.get("SPE_TEST_GRAPH", "0") == "1"
    cfg = O.tiny_config()
    params = O.make_params(cfg, 51)
    Bg = 4 * world
    images, targets = O.make_inputs(cfg, Bg, 48, 64, seed=51, max_gt=4)
    tg = [{k: v.to(dev) for k, v in t.items()} for t in targets]
    tr = [dict(t, scores=torch.full((len(t["labels"]),), 0.7, device=dev)) for t in tg]

    def build():
        model = factory.build_detector(cfg, dev).train()
        model.load_state_dict(params)
        crit = factory.build_criterion(cfg, device=dev).eval()
        crit_r = factory.build_criterion(cfg, refine=True, device=dev).eval()
        return model, TrainStep(model, crit, crit_r, graph=graph, max_gt=8)

    # single-process reference on the full batch (before the process group exists: num_boxes is not all-reduced)
    model0, step0 = build()
    loss0, ld0, _ = step0(images.to(dev), tg, tr)
    ref_flat = step0.gbuf.flat.clone()
    ref_loss = float(loss0)
    ref_ld = {k: float(v) for k, v in ld0.items()}
    del model0, step0

    dist.init_process_group("nccl", device_id=torch.device("cuda", torch.cuda.current_device()))
    model, step = build()
    sl = slice(rank * 4, (rank + 1) * 4)
    loss, ld, _ = step(images[sl].to(dev), tg[sl], tr[sl])
    flat = step.gbuf.flat
    # loss: mean over ranks == full-batch loss (the num_boxes normaliser is all-reduced / world, conditional_detr.py:436-440)
    l = loss.detach().clone().reshape(1)
    dist.all_reduce(l, op=dist.ReduceOp.AVG)
    err_loss = abs(float(l) - ref_loss) / abs(ref_loss)
    num = float((flat - ref_flat).norm())
    den = float(ref_flat.norm())
    # every rank holds the same reduced buffer
    chk = flat.clone()
    dist.all_reduce(chk, op=dist.ReduceOp.MAX)
    same = float((chk - flat).abs().max())
    ok = err_loss < 2e-3 and num / den < 5e-3 and same == 0.0
    print("rank %d loss %.6f (avg %.6f vs %.6f, rel %.2e) grad rel err %.2e identical-across-ranks %s -> %s"
          % (rank, float(loss), float(l), ref_loss, err_loss, num / den, same == 0.0, "OK" if ok else "FAIL"), flush=True)
    # second step (graph mode: a pure replay, the bucketed all-reduce captured inside it) must land on the same gradients
    step(images[sl].to(dev), tg[sl], tr[sl])
    err2 = float((step.gbuf.flat - ref_flat).norm()) / den
    if err2 > 5e-3:
        ok = False
    print("rank %d second step grad rel err %.2e  bucketed all-reduce split at %s" % (rank, err2, step._split), flush=True)
    for k in ("loss_ce", "loss_bbox", "loss_giou"):
        v = ld[k].detach().clone().reshape(1)
        dist.all_reduce(v, op=dist.ReduceOp.AVG)
        if abs(float(v) - ref_ld[k]) > 2e-3 * abs(ref_ld[k]) + 1e-5:
            ok = False
            print("rank %d %s avg %.6f vs %.6f FAIL" % (rank, k, float(v), ref_ld[k]), flush=True)
    dist.barrier()
    step.close()            # captured collectives keep the communicator (and destroy_process_group) waiting
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
