"""CPU, gloo, world_size 2: the data-parallel host logic -- flat gradient buffer + single all-reduce, and the criterion's
global num_boxes normaliser (conditional_detr.py:436-440)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from spe_b200.dp import FlatGradBuffer
        from spe_b200.criterion_ops import PackedTargets
        torch.manual_seed(0)
        m = torch.nn.Sequential(torch.nn.Linear(5, 7), torch.nn.Linear(7, 3))
        buf = FlatGradBuffer(m.parameters())
        x = torch.full((4, 5), float(rank + 1))
        m(x).sum().backward()                      # autograd accumulates INTO the flat views
        local = buf.flat.clone()
        assert all(p.grad.data_ptr() >= buf.flat.data_ptr() for p in m.parameters())
        buf.all_reduce_mean()
        gathered = [torch.zeros_like(local) for _ in range(world)]
        dist.all_gather(gathered, local)
        torch.testing.assert_close(buf.flat, sum(gathered) / world)
        # num_boxes: rank r has r+1 boxes -> mean 1.5 -> inv 1/1.5
        tg = [{"labels": torch.ones(rank + 1, dtype=torch.int64), "boxes": torch.rand(rank + 1, 4)}]
        T = PackedTargets(tg, torch.device("cpu"))
        torch.testing.assert_close(T.inv_num_boxes(), torch.tensor([1.0 / 1.5]))
        q.put((rank, "ok"))
    except Exception as e:  # pragma: no cover
        q.put((rank, repr(e)))
    finally:
        dist.destroy_process_group()


def test_flat_grad_allreduce_and_num_boxes_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(60)
    assert sorted(res) == [(0, "ok"), (1, "ok")], res
