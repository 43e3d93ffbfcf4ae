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


def _model_worker(rank, world, port, ret):
    """The REAL model on two gloo ranks (host tensors -> the library's CPU backend): windows shard over the batch, the
    flat-bucket all-reduce averages the gradients, and the result equals the mean of the per-shard oracle gradients."""
    import numpy as np
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(2)
    import mtad_gat_pytorch_b200 as mg
    from mtad_gat_pytorch_b200.training import allreduce_gradients, shard_batch, rmse_losses
    from oracle import mtad_gat_oracle as orc
    kw = dict(n_features=6, window_size=16, out_dim=6, kernel_size=3, gru_hid_dim=12, forecast_n_layers=2,
              forecast_hid_dim=10, recon_hid_dim=9, dropout=0.0)
    cfg = orc.Config(**kw)
    params = orc.make_params(cfg, seed=3, dtype=np.float64)
    rng = np.random.default_rng(3)
    Bg = 10
    X, Y = rng.random((Bg, cfg.n, cfg.k)), rng.random((Bg, 1, cfg.k))
    m = mg.MTAD_GAT(**kw)
    m.load_state_dict({k: torch.from_numpy(v.astype(np.float32)) for k, v in params.items()})
    m.train()
    lo, hi = shard_batch(Bg, world, rank)
    x = torch.from_numpy(X[lo:hi].astype(np.float32)); y = torch.from_numpy(Y[lo:hi].astype(np.float32))
    preds, recons = m(x)
    fl, rl = rmse_losses(x, y, preds, recons)
    (fl + rl).backward()
    plist = [p for p in m.parameters()]
    allreduce_gradients(plist, world)
    g_mean = None
    for r in range(world):
        a, b = shard_batch(Bg, world, r)
        g = orc.loss_fwd_bwd(X[a:b], Y[a:b], params, cfg)[6]
        g_mean = g if g_mean is None else {k: g_mean[k] + g[k] for k in g}
    err = 0.0
    for name, p in m.named_parameters():
        ref = g_mean[name] / world
        err = max(err, float(np.abs(p.grad.numpy() - ref).max() / max(np.abs(ref).max(), 1e-9)))
    ret[rank] = err
    dist.destroy_process_group()


def test_real_model_gradients_average_across_two_gloo_ranks():
    import __graft_entry__ as ge
    ge.build()
    mgr = mp.Manager()
    ret = mgr.dict()
    port = 29100 + os.getpid() % 1500
    mp.spawn(_model_worker, args=(2, port, ret), nprocs=2, join=True)
    assert len(ret) == 2 and all(v < 2e-4 for v in ret.values()), dict(ret)
