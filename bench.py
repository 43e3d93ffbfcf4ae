#!/usr/bin/env python
"""bench.py -- windows/sec of the MTAD-GAT training step (forward + loss + backward + Adam) at SMD shape.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--no-graph] [--batch B]

Workload (BASELINE.json configs[1]): MTAD_GAT(k=38, n=100, out=38, forecast_n_layers=3, dropout=0.3) in train
mode, batch 256 windows per GPU of synthetic uniform[0,1) data, loss = sqrt(mse)+sqrt(mse) as training.py:122-124,
torch.optim.Adam(lr=1e-3).  One "step" = one pass of the hot path over one batch, optimizer step included.

JSON line keys follow the driver contract: value (device-resident inputs, CUDA-event timed, L2 flushed between
steps), e2e (host pinned inputs -> H2D -> step -> D2H loss, wall clock), roofline (per-kernel, measured live with
CUDA events), cpu_baseline (the numpy oracle port timed on a bounded sample on the host cores), clocks.
`--impl reference` times the CPU port of the reference path (oracle/) instead -- the reference is Python/torch
and cannot travel to the GPU box, see DESIGN.md.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

K_FEAT, N_WIN, OUT_DIM, BATCH = 38, 100, 38, 256
MODEL_KW = dict(n_features=K_FEAT, window_size=N_WIN, out_dim=OUT_DIM, forecast_n_layers=3, dropout=0.3)
WORKLOAD = "SMD-shape (k=38,n=100) MTAD_GAT train step fwd+bwd+Adam, batch 256/GPU, fp32"


# dram__bytes_read.sum + dram__bytes_write.sum per launch at the default workload, from the `ncu --set full` capture
# summarised in profiles/ (cold caches: ncu flushes L2 before every replay pass)
NCU_TRAFFIC = {"gru_recurrence_bwd_kernel": 108.3e6, "gru_recurrence_fwd_kernel": 73.3e6}   # profiles/r1_final_gru_cl_raw.csv


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return {"hbm_gbs": d["hbm_gbs"], "bf16_tflops": d.get("bf16_tflops_sustained", d["bf16_tflops"]), "src": "measured"}
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1400.0, "src": "fallback"}


# ----------------------------------------------------------------------------------------------------------
# CPU baseline / reference arm: the oracle port on the host cores
# ----------------------------------------------------------------------------------------------------------
def cpu_port_rate(sample_b, reps, seed=0):
    """windows/s of the numpy port of the reference path (forward as written + backward), fp32, `sample_b`
    windows per pass.  Adam is excluded (it is <1% of the CPU time)."""
    from oracle import mtad_gat_oracle as orc
    cfg = orc.Config(**MODEL_KW)
    params = orc.make_params(cfg, seed=seed, dtype=np.float32)
    rng = np.random.default_rng(seed)
    x = rng.random((sample_b, N_WIN, K_FEAT)).astype(np.float32)
    y = rng.random((sample_b, 1, K_FEAT)).astype(np.float32)
    orc.loss_fwd_bwd(x, y, params, cfg)            # warm-up (BLAS threads, page faults)
    times = []
    for _ in range(reps):
        t0 = time.perf_counter()
        orc.loss_fwd_bwd(x, y, params, cfg)
        times.append(time.perf_counter() - t0)
    return sample_b / min(times), sample_b / (sum(times) / len(times)), times


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    sample_b = 16
    cores = os.cpu_count()
    # warmup passes
    for _ in range(max(0, args.warmup - 1)):
        cpu_port_rate(sample_b, 1)
    best, mean, times = cpu_port_rate(sample_b, max(1, args.steps))
    ms = 1e3 * sum(times) / len(times)
    line = {
        "impl": "reference", "metric": "windows/sec MTAD-GAT fwd+bwd (k=38,n=100)", "value": mean, "unit": "windows/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "fp32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "sample": f"{sample_b} windows per step on the CPU"},
        "cpu_baseline": {"value": mean, "unit": "windows/s", "cores": cores, "kind": "port",
                         "sample": f"{sample_b}-window fwd+bwd passes of the numpy port (oracle/), {len(times)} timed"},
        "e2e": {"value": mean, "unit": "windows/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------------------------------------
class ClockSampler:
    """SM clock and throttle reasons sampled DURING the timed region: NVML (the library behind nvidia-smi; ~1 ms per
    sample) when importable, else the `nvidia-smi --query-gpu=clocks.sm,...` line of the profiling recipe."""
    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")
    NAMES = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]

    def __init__(self, gpu_index):
        self.idx = gpu_index
        self.samples = []          # (sm_mhz, [reason flags])
        self.sm_max = None
        self.source = None
        self.stop = threading.Event()
        self.th = None
        self.nv = None
        try:
            import pynvml
            pynvml.nvmlInit()
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = int(vis.split(",")[gpu_index]) if vis and all(v.strip().isdigit() for v in vis.split(",")) else gpu_index
            self.h = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.sm_max = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
            self.nv = pynvml
            self.source = "nvml"
        except Exception:
            self.nv = None
            self.source = "nvidia-smi"

    def _sample_nvml(self):
        nv = self.nv
        sm = float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
        r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
        flags = [bool(r & nv.nvmlClocksEventReasonHwSlowdown), bool(r & nv.nvmlClocksEventReasonHwThermalSlowdown),
                 bool(r & nv.nvmlClocksEventReasonSwThermalSlowdown), bool(r & nv.nvmlClocksEventReasonSwPowerCap)]
        self.samples.append((sm, flags))

    def _sample_smi(self):
        out = subprocess.run(["nvidia-smi", "-i", str(self.idx), f"--query-gpu={self.FIELDS}",
                              "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
        parts = [p.strip() for p in out.strip().split(",")]
        if len(parts) >= 7 and parts[0].replace(".", "").isdigit():
            self.sm_max = float(parts[1])
            self.samples.append((float(parts[0]), [parts[3 + i].lower().startswith("active") for i in range(4)]))

    def _run(self):
        while not self.stop.is_set():
            try:
                if self.nv is not None:
                    self._sample_nvml()
                else:
                    self._sample_smi()
            except Exception:
                pass
            self.stop.wait(0.005 if self.nv is not None else 0.05)

    def __enter__(self):
        self.th = threading.Thread(target=self._run, daemon=True)
        self.th.start()
        return self

    def __exit__(self, *a):
        self.stop.set()
        self.th.join(timeout=6)

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.sm_max, "reasons": ["clock sampling unavailable"], "samples": 0}
        sm = sorted(s[0] for s in self.samples)
        reasons = [n for i, n in enumerate(self.NAMES) if any(s[1][i] for s in self.samples)]
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": self.sm_max, "reasons": reasons, "samples": len(self.samples),
                "source": self.source}


def run_ours(args):
    import torch
    import torch.distributed as dist
    import mtad_gat_pytorch_b200 as mg
    from mtad_gat_pytorch_b200 import training as mgt

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    B = args.batch
    if args.gru_split:
        mg.set_gru_split(args.gru_split)
    torch.manual_seed(0)
    model = mg.MTAD_GAT(**MODEL_KW).to(dev)
    model.train()
    opt = torch.optim.Adam(model.parameters(), lr=1e-3, capturable=not args.no_graph, fused=True)
    step = mgt.TrainStep(model, opt, batch=B, use_graph=not args.no_graph, world_size=world)

    g = torch.Generator(device="cpu").manual_seed(1234 + rank)
    n_host = 8
    xs_host = [torch.rand(B, N_WIN, K_FEAT, generator=g).pin_memory() for _ in range(n_host)]
    ys_host = [torch.rand(B, 1, K_FEAT, generator=g).pin_memory() for _ in range(n_host)]
    xs_dev = [t.to(dev) for t in xs_host]
    ys_dev = [t.to(dev) for t in ys_host]
    flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device=dev)   # > 126 MB L2

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- warm-up (also captures the CUDA graph) ----
    for i in range(max(args.warmup, 3)):
        step.run_device(xs_dev[i % n_host], ys_dev[i % n_host])
    barrier()

    # ---- timed region: device-resident inputs, per-step CUDA events, L2 flushed between steps ----
    mg.reset_launch_count()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    with ClockSampler(local) as clocks:
        barrier()
        t_wall0 = time.perf_counter()
        for i in range(args.steps):
            flush.zero_()
            evs[i][0].record()
            step.run_device(xs_dev[i % n_host], ys_dev[i % n_host])
            evs[i][1].record()
        barrier()
        t_wall = time.perf_counter() - t_wall0
        ms_dev = sum(a.elapsed_time(b) for a, b in evs) / args.steps
        launches = step.launches_per_step * args.steps if step.launches_per_step else mg.launch_count()

        # ---- e2e: host pinned inputs -> H2D -> step -> D2H loss, wall clock over K steps ----
        barrier()
        t0 = time.perf_counter()
        last = None
        # software pipeline of the training loop: H2D of batch i+1 (copy stream) and the host-side enqueue of step
        # i+1 overlap step i; every step still copies its batch in and its loss out (the loss of step i is read on
        # the host after step i+1 has been enqueued)
        step.prefetch_host(xs_host[0], ys_host[0])
        for i in range(args.steps):
            if i + 1 < args.steps:
                step.prefetch_host(xs_host[(i + 1) % n_host], ys_host[(i + 1) % n_host])
            step.launch_prefetched()
            if i > 0:
                last = step.collect()
        last = step.collect()
        torch.cuda.synchronize()
        e2e_s = (time.perf_counter() - t0) / args.steps     # local wall time; MAX over ranks below (the closing
        barrier()                                           # collective barrier itself is not part of the K steps)
    t_ms = torch.tensor([ms_dev, e2e_s * 1e3], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t_ms, op=dist.ReduceOp.MAX)
    ms_dev, e2e_ms = float(t_ms[0]), float(t_ms[1])

    if rank == 0:
        peaks = load_peaks()
        from mtad_gat_pytorch_b200 import kernel_bench
        kern = kernel_bench.stage_rooflines(model, B, peaks, dev)
        # dominant single kernel = the slower of the two persistent recurrence launches (largest share of the step,
        # profiles/r1_launches_bench.csv); the other rows are whole stages (several launches each)
        single = [r for r in kern if r.get("single_kernel")]
        dominant = max(single or kern, key=lambda r: r["ms"])
        if B == BATCH and dominant["kernel"] in NCU_TRAFFIC:
            dominant["traffic"] = NCU_TRAFFIC[dominant["kernel"]]
        roof = {k_: dominant[k_] for k_ in ("bound", "achieved", "peak", "unit", "frac", "traffic")}
        roof["kernel"] = dominant["kernel"]
        roof["peak_src"] = peaks["src"]
        cpu = None
        if world == 1 and not args.skip_cpu:
            best, mean, times = cpu_port_rate(16, 3)
            cpu = {"value": mean, "unit": "windows/s", "cores": os.cpu_count(), "kind": "port",
                   "sample": f"3 timed 16-window fwd+bwd passes of the numpy port of the reference path ({sum(times):.1f} s)"}
        line = {
            "metric": "windows/sec MTAD-GAT fwd+bwd (k=38,n=100)", "value": B * world / (ms_dev * 1e-3), "unit": "windows/s",
            "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms_dev,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "fp32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "global_batch": B * world, "parallelism": f"dp{world}",
                       "e2e_loop": "pinned host batches, H2D of batch i+1 and enqueue of step i+1 overlap step i, loss of every step copied back",
                       "cuda_graph": not args.no_graph, "l2": "256 MiB buffer zeroed between timed steps",
                       "optimizer": "torch.optim.Adam(fused) inside the step"},
            "e2e": {"value": B * world / (e2e_ms * 1e-3), "unit": "windows/s",
                    "h2d_bytes_per_step": int(xs_host[0].numel() * 4 + ys_host[0].numel() * 4), "d2h_bytes_per_step": 8,
                    "ms_per_step": e2e_ms},
            "gpu_launches": int(launches), "roofline": roof, "kernels": kern, "cpu_baseline": cpu,
            "clocks": clocks.summary(), "wall_s_timed": t_wall, "loss": last,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=BATCH)
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--skip-cpu", action="store_true")
    ap.add_argument("--gru-split", type=int, default=0, help="clusters per 16-window tile in the recurrence (0 = auto)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
