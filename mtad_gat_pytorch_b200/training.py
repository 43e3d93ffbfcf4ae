"""Training-step driver for the B200 path: the reference's step (training.py:106-127: zero_grad, forward,
sqrt-MSE losses, backward, Adam) with static device buffers, optionally captured once in a CUDA graph and
replayed (the step is launch-bound at batch 256), and with a flat-bucket NCCL gradient all-reduce when
world_size > 1 (one process per GPU; windows shard over the batch, parameters are replicated)."""
import torch
import torch.distributed as dist

from . import functional as F


def rmse_losses(x, y, preds, recons, target_dims=None):
    """training.py:113-124.  CUDA tensors of matching element counts take the fused two-launch kernel
    (mtadgat_rmse_pair_fwd / _bwd); CPU tensors (host-side tests) use the plain expression."""
    if target_dims is not None:
        x = x[:, :, target_dims]
        y = y[:, :, target_dims].squeeze(-1)
    if preds.ndim == 3:
        preds = preds.squeeze(1)
    if y.ndim == 3:
        y = y.squeeze(1)
    if preds.is_cuda and preds.shape == y.shape and recons.shape == x.shape and preds.dtype == torch.float32:
        return F.RmsePairFn.apply(preds, y, recons, x)
    fl = torch.sqrt(torch.mean((y - preds) ** 2))
    rl = torch.sqrt(torch.mean((x - recons) ** 2))
    return fl, rl


def allreduce_gradients(params, world_size):
    """Average the gradients of `params` across ranks with ONE flat-bucket all-reduce (1.6 MB at SMD shape).
    Backend-agnostic (NCCL on the GPUs, gloo in the CPU tests)."""
    if world_size == 1:
        return
    grads = [p.grad for p in params]
    flat = torch._utils._flatten_dense_tensors(grads)
    if dist.get_backend() == "nccl":
        dist.all_reduce(flat, op=dist.ReduceOp.AVG)        # the average is formed inside the collective
    else:
        dist.all_reduce(flat)
        flat.div_(world_size)
    torch._foreach_copy_(grads, torch._utils._unflatten_dense_tensors(flat, grads))


def shard_batch(global_batch, world_size, rank):
    """Windows shard over the batch: contiguous [lo, hi) slice of the global batch owned by `rank`."""
    base, rem = divmod(global_batch, world_size)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


class TrainStep:
    def __init__(self, model, optimizer, batch, use_graph=True, world_size=1, target_dims=None):
        p0 = next(model.parameters())
        self.model, self.opt, self.world = model, optimizer, world_size
        self.target_dims = target_dims
        n, k = model.temporal_gat.window_size, model.temporal_gat.n_features
        self.x = torch.zeros(batch, n, k, device=p0.device)
        self.y = torch.zeros(batch, 1, k, device=p0.device)
        self.losses = torch.zeros(2, device=p0.device)
        self.params = [p for p in model.parameters() if p.requires_grad]
        self.use_graph = use_graph
        # host-input pipeline: two device staging buffers filled by a copy stream, so the H2D copy of batch i+1
        # overlaps the step on batch i (every batch is still copied inside the timed region)
        self._stage = [(torch.zeros_like(self.x), torch.zeros_like(self.y)) for _ in range(2)]
        self._copy_stream = torch.cuda.Stream(device=p0.device)
        self._staged = [None, None]
        self._put = self._get = 0
        self.g_fb = self.g_opt = None
        self.launches_per_step = 0
        self._warm = 0

    # -- pieces ------------------------------------------------------------------------------------------
    def _fwd_bwd(self):
        self.opt.zero_grad(set_to_none=True)
        preds, recons = self.model(self.x)
        fl, rl = rmse_losses(self.x, self.y, preds, recons, self.target_dims)
        (fl + rl).backward()
        self.losses[0].copy_(fl.detach())
        self.losses[1].copy_(rl.detach())

    def _allreduce(self):
        allreduce_gradients(self.params, self.world)

    def _capture(self):
        # warm-up and capture on the SAME stream: the library's GEMM pack workspaces are per stream and cannot grow
        # during capture (include/mtadgat.h: mtadgat_workspace_reserve)
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            for _ in range(3):
                self._fwd_bwd(); self._allreduce(); self.opt.step()
        torch.cuda.current_stream().wait_stream(s)
        torch.cuda.synchronize()
        F.reset_launch_count()
        if self.world == 1:
            self.g_fb = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self.g_fb, stream=s):
                self._fwd_bwd()
                self.opt.step()
        else:
            self.g_fb = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self.g_fb, stream=s):
                self._fwd_bwd()
            self.g_opt = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self.g_opt, stream=s):
                self.opt.step()
        self.launches_per_step = F.launch_count()

    def _run(self):
        if not self.use_graph:
            self._fwd_bwd(); self._allreduce(); self.opt.step()
            return
        if self.g_fb is None:
            self._capture()
        self.g_fb.replay()
        if self.world > 1:
            self._allreduce()
            self.g_opt.replay()

    # -- public ------------------------------------------------------------------------------------------
    def run_device(self, x, y):
        """x (B,n,k), y (B,1,k) already on the device."""
        self.x.copy_(x); self.y.copy_(y)
        self._run()

    def prefetch_host(self, x_host, y_host):
        """Start the H2D copy of a pinned host batch on the copy stream (call before run_prefetched)."""
        slot = self._put & 1
        sx, sy = self._stage[slot]
        cs = self._copy_stream
        cs.wait_stream(torch.cuda.current_stream())      # the step that last read this slot has been enqueued
        with torch.cuda.stream(cs):
            sx.copy_(x_host, non_blocking=True); sy.copy_(y_host, non_blocking=True)
            ev = torch.cuda.Event(); ev.record(cs)
        self._staged[slot] = ev
        self._put += 1

    def run_prefetched(self):
        """Run the step on the oldest prefetched batch; returns the scalar loss (D2H)."""
        slot = self._get & 1
        self._get += 1
        torch.cuda.current_stream().wait_event(self._staged[slot])
        sx, sy = self._stage[slot]
        self.x.copy_(sx); self.y.copy_(sy)
        self._run()
        fl, rl = self.losses.tolist()
        return fl + rl

    def launch_prefetched(self):
        """Enqueue the step on the oldest prefetched batch plus an asynchronous D2H copy of its two loss terms into
        pinned host memory, without waiting: the host can enqueue step i+1 while step i runs.  collect() returns the
        losses in launch order."""
        slot = self._get & 1
        self._get += 1
        torch.cuda.current_stream().wait_event(self._staged[slot])
        sx, sy = self._stage[slot]
        self.x.copy_(sx); self.y.copy_(sy)
        self._run()
        if not hasattr(self, "_loss_host"):
            self._loss_host = [torch.zeros(2, dtype=torch.float32).pin_memory() for _ in range(2)]
            self._loss_ev = [None, None]
            self._lput = self._lget = 0
        k = self._lput & 1
        self._lput += 1
        self._loss_host[k].copy_(self.losses, non_blocking=True)
        ev = torch.cuda.Event(); ev.record()
        self._loss_ev[k] = ev

    def collect(self):
        """Scalar loss of the oldest launched-but-uncollected step (waits for that step only)."""
        k = self._lget & 1
        self._lget += 1
        self._loss_ev[k].synchronize()
        v = self._loss_host[k]
        return float(v[0]) + float(v[1])

    def run_host(self, x_host, y_host):
        """Pinned host batch in, scalar loss out (H2D + step + D2H)."""
        self.x.copy_(x_host, non_blocking=True); self.y.copy_(y_host, non_blocking=True)
        self._run()
        fl, rl = self.losses.tolist()
        return fl + rl
