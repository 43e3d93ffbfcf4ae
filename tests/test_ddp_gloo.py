"""world_size-2 gloo test (CPU) of the data-parallel logic: windows shard over the batch, one flat-bucket
all-reduce averages the gradients, replicas stay identical."""
import os

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _worker(rank, world, port, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from mtad_gat_pytorch_b200.training import allreduce_gradients, shard_batch
    torch.manual_seed(0)
    net = torch.nn.Sequential(torch.nn.Linear(6, 5), torch.nn.Tanh(), torch.nn.Linear(5, 3))
    X = torch.randn(10, 6, generator=torch.Generator().manual_seed(1))
    lo, hi = shard_batch(10, world, rank)
    loss = net(X[lo:hi]).pow(2).sum() / 10.0          # global-mean loss split over shards ...
    loss.backward()
    params = list(net.parameters())
    allreduce_gradients(params, world)                 # ... averaged, so multiply back by world to compare
    g = torch.cat([p.grad.reshape(-1) for p in params]) * world
    net2 = torch.nn.Sequential(torch.nn.Linear(6, 5), torch.nn.Tanh(), torch.nn.Linear(5, 3))
    net2.load_state_dict(net.state_dict())
    (net2(X).pow(2).sum() / 10.0).backward()
    g_full = torch.cat([p.grad.reshape(-1) for p in net2.parameters()])
    ret[rank] = float((g - g_full).abs().max())
    dist.destroy_process_group()


def test_flat_bucket_allreduce_world2():
    mgr = mp.Manager()
    ret = mgr.dict()
    port = 29000 + os.getpid() % 2000
    mp.spawn(_worker, args=(2, port, ret), nprocs=2, join=True)
    assert len(ret) == 2 and all(v < 1e-6 for v in ret.values()), dict(ret)
