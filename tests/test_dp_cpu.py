"""world_size-2 gloo test of the data-parallel step's single flat all-reduce (CPU, no kernels)."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from graph_neural_net_b200.training import flat_gradient_allreduce, shard_bounds


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.manual_seed(0)
    model = torch.nn.Sequential(torch.nn.Linear(6, 5), torch.nn.ReLU(), torch.nn.Linear(5, 3))
    data = torch.arange(8 * 6, dtype=torch.float32).reshape(8, 6) / 10.0
    lo, hi = shard_bounds(8, rank, world)
    loss_sum = model(data[lo:hi]).pow(2).sum()              # un-normalised local objective
    loss_sum.backward()
    extras = torch.tensor([float(loss_sum), float(hi - lo)])
    red = flat_gradient_allreduce(model.parameters(), extras)
    grads = torch.cat([p.grad.reshape(-1) for p in model.parameters()]) / red[1]
    out[rank] = (grads.clone(), red.clone())
    dist.destroy_process_group()


def test_flat_allreduce_matches_single_process():
    world = 2
    port = _free_port()
    with mp.Manager() as mgr:
        out = mgr.dict()
        mp.spawn(_worker, args=(world, port, out), nprocs=world, join=True)
        results = {k: v for k, v in out.items()}
    torch.manual_seed(0)
    model = torch.nn.Sequential(torch.nn.Linear(6, 5), torch.nn.ReLU(), torch.nn.Linear(5, 3))
    data = torch.arange(8 * 6, dtype=torch.float32).reshape(8, 6) / 10.0
    total = model(data).pow(2).sum()
    (total / 8).backward()
    ref = torch.cat([p.grad.reshape(-1) for p in model.parameters()])
    for rank in range(world):
        grads, red = results[rank]
        assert torch.allclose(grads, ref, rtol=1e-5, atol=1e-6)
        assert abs(float(red[0]) - float(total)) < 1e-3 * abs(float(total)) and float(red[1]) == 8.0


def test_shard_bounds_cover_everything():
    for total in (1, 7, 64, 1024):
        for world in (1, 2, 3, 8):
            spans = [shard_bounds(total, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            assert max(hi - lo for lo, hi in spans) - min(hi - lo for lo, hi in spans) <= 1
