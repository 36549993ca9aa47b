#!/usr/bin/env python
"""
bench.py -- grid-cells*steps/s of the U-RNN encoder-decoder time step (flood-depth forward) on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--math f16x3|fp32|bf16]
                    [--scaling weak|strong] [--height H --width W]

One "step" = one ED time step (reference model.py:65-121) over the whole H x W grid.  Workload at N=1: the
location1 full-resolution grid (500 x 500, 2 m / 1 min, historical_nums=30 => C_in=63), synthetic event,
seeded random-init weights (no dataset / checkpoint offline).  Prints ONE JSON line.

  value     : cells*steps/s with the per-step inputs already resident in HBM: urnn_ed_sequence_dev over chunks of distinct
              dense (C_in,H,W) inputs (default math f16x3: tcgen05, fp16 hi+lo split operands -- the fast mode that passes
              the T=180 parity gate, tests/test_gpu_x3.py)
  e2e       : the reference's own workflow (test.py:447-375) through the C ABI with HOST buffers, urnn_ed_event_host: the
              event (3 static maps + scalar rainfall series) is uploaded once, every step's depth map returns to the host
              inside the timed region.  e2e_dense: urnn_ed_sequence_host, one (C_in,H,W) H2D copy per step.
  roofline  : the decoder stage-1 ConvGRU cell (dominant) and the encoder stage-1 cell (worst), per-launch CUDA-event times
              measured inside the running step (urnn_ed_profile_dev), algorithmic bytes vs MEASURED_PEAKS.json hbm_gbs
  cpu_baseline / --impl reference : the UNMODIFIED reference modules (oracle/_ref, copied from /root/reference by
              oracle/make_ref.py) on this box's host cores; falls back to the op-for-op torch port if the copy is absent
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, "u-rnn_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)

import numpy as np  # noqa: E402
import torch  # noqa: E402

METRIC = "grid-cells*steps/s (flood-depth fwd)"
UNIT = "cells*steps/s"
H_DEF, W_DEF, HIST_DEF = 500, 500, 30


def trace(msg):
    if os.environ.get("URNN_BENCH_TRACE"):
        print(f"[bench {time.strftime('%H:%M:%S')}] {msg}", file=sys.stderr, flush=True)


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=180)   # T=180: BASELINE config 3
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--math", default=os.environ.get("URNN_MATH", "f16x3"), choices=["fp32", "bf16", "f16x3"])
    ap.add_argument("--mode", default="infer", choices=["infer", "train"],
                    help="train: BASELINE config 4 -- one step = one SWP window (seq_num time steps forward with autograd, backward, "
                         "gradient all-reduce, clip, Adam), fp32 kernels")
    ap.add_argument("--seq-num", type=int, default=12, help="--mode train: time steps per window (location1_scratch.yaml:56)")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="N > 1: weak = every rank owns an H x W band of an (N*H) x W grid; strong = the H x W grid is split into N bands")
    ap.add_argument("--height", type=int, default=H_DEF)
    ap.add_argument("--width", type=int, default=W_DEF)
    ap.add_argument("--hist", type=int, default=HIST_DEF)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--watchdog", type=int, default=900, help="seconds before a stuck run dumps its stack and exits")
    ap.add_argument("--value-only", action="store_true", help="device-resident throughput only (large grids: skips the e2e / roofline legs)")
    return ap.parse_args()


# ---------------------------------------------------------------------------------------------- helpers
def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """SM clocks / throttle reasons sampled (every 50 ms) while the timed region runs (B200_PROFILING.md's clocks line).
    NVML in-process (the library nvidia-smi itself reads; initialised before the timed region): an `nvidia-smi -lms` poller
    holds driver locks for tens of milliseconds per sample, which a multi-GPU run pays at every cross-GPU exchange (measured:
    8-GPU weak scaling 2.26 ms/step with the poller vs 1.02 ms/step without).  Falls back to nvidia-smi where pynvml is absent."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
    PERIOD = 0.05          # NVML queries are cheap (tens of microseconds); the nvidia-smi fallback polls every 200 ms

    def __init__(self, index=0, enabled=True):
        self.index, self.proc, self.lines, self.enabled = index, None, [], enabled
        self.nvml, self.handle, self.stop, self.thread, self.samples = None, None, threading.Event(), None, []
        if not enabled:
            return
        try:
            import pynvml
            pynvml.nvmlInit()
            self.handle = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.handle, pynvml.NVML_CLOCK_SM))
            self.nvml = pynvml
        except Exception:
            self.nvml = None

    def _poll_nvml(self):
        n = self.nvml
        bits = [("hw_slowdown", n.nvmlClocksEventReasonHwSlowdown if hasattr(n, "nvmlClocksEventReasonHwSlowdown") else n.nvmlClocksThrottleReasonHwSlowdown),
                ("hw_thermal_slowdown", getattr(n, "nvmlClocksEventReasonHwThermalSlowdown", None) or n.nvmlClocksThrottleReasonHwThermalSlowdown),
                ("sw_thermal_slowdown", getattr(n, "nvmlClocksEventReasonSwThermalSlowdown", None) or n.nvmlClocksThrottleReasonSwThermalSlowdown),
                ("sw_power_cap", getattr(n, "nvmlClocksEventReasonSwPowerCap", None) or n.nvmlClocksThrottleReasonSwPowerCap)]
        get = getattr(n, "nvmlDeviceGetCurrentClocksEventReasons", None) or n.nvmlDeviceGetCurrentClocksThrottleReasons
        while True:
            try:
                mhz = float(n.nvmlDeviceGetClockInfo(self.handle, n.NVML_CLOCK_SM))
                mask = int(get(self.handle))
                self.samples.append((mhz, [nm for nm, b in bits if mask & int(b)]))
            except Exception:
                pass
            if self.stop.wait(self.PERIOD):
                break

    def __enter__(self):
        if not self.enabled:       # multi-rank runs: one sampler (rank 0)
            return self
        if self.nvml is not None:
            self.thread = threading.Thread(target=self._poll_nvml, daemon=True)
            self.thread.start()
            return self
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None
        return self

    def _read(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def __exit__(self, *a):
        self.stop.set()
        if self.nvml is not None and self.thread is not None:
            self.thread.join(timeout=2)
        if self.proc is not None:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()

    def summary(self):
        sm, mx, reasons = [], 0, set()
        if self.nvml is not None:
            for mhz, rs in self.samples:
                sm.append(mhz); reasons.update(rs)
            return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": self.max_mhz if sm else None,
                    "reasons": sorted(reasons), "samples": len(sm), "source": "nvml"}
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [t.strip() for t in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx = max(mx, float(f[1]))
            except ValueError:
                continue
            for n, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx or None,
                "reasons": sorted(reasons), "samples": len(sm), "source": "nvidia-smi"}


def build_net(H, W, C, math, device):
    from src.lib.model.networks.model import ED
    from src.lib.model.networks.net_params import get_network_params
    torch.manual_seed(0)
    enc, dec = get_network_params(False, H, W, input_channels=C, math=math)
    return ED(False, enc, dec, 0.5, False, input_height=H, input_width=W).to(device).eval()


def synthetic_inputs(H, W, hist, T, rain_scale=6.0, rain_max=6.0):
    from oracle.urnn_oracle import synthetic_event_inputs   # input generator only (numpy recipe)
    return synthetic_event_inputs(H, W, T, hist, seed=42, rain_scale=rain_scale, rain_max=rain_max)


def state_shapes(H, W):
    return [(1, 64, H, W), (1, 96, H // 2, W // 2), (1, 96, H // 4, W // 4),
            (1, 96, H // 4, W // 4), (1, 96, H // 2, W // 2), (1, 64, H, W)]


# ---------------------------------------------------------------------------------------------- CPU arm
def workload_name(H, W, C, world=1, scaling="weak"):
    grid = f"{H}x{W}" if world == 1 else (f"{world * H}x{W} ({world} row bands of {H}x{W})" if scaling == "weak"
                                          else f"{H}x{W} ({world} row bands of {H // world}x{W})")
    return f"location1 full-res {grid}, C_in={C}, ED step forward (6 ConvGRU cells + stems + head)"


def cpu_reference_throughput(H, W, hist, steps, warmup):
    """The reference's own modules (oracle/_ref: unmodified copies made by oracle/make_ref.py) on the host cores, no_grad,
    use_checkpoint=False, all threads; the op-for-op torch port (oracle/torch_port.py) if the copy is not there."""
    C = 2 * hist + 3
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    xs = torch.from_numpy(synthetic_inputs(H, W, hist, 2))
    ref_dir = os.path.join(ROOT, "oracle", "_ref")
    if os.path.exists(os.path.join(ref_dir, "src", "lib", "model", "networks", "model.py")):
        kind = "reference"
        for m in [k for k in sys.modules if k == "src" or k.startswith("src.")]:
            del sys.modules[m]
        saved = list(sys.path)
        sys.path[:] = [ref_dir] + [p for p in saved if not p.rstrip("/").endswith("u-rnn_b200")]
        torch.Tensor.cuda = lambda self, *a, **k: self            # hard-coded .cuda() in ConvRNN.py:136,146
        try:
            from src.lib.model.networks.net_params import get_network_params
            from src.lib.model.networks.model import ED
            from src.lib.utils.net_config import load_net_config
            cfg = load_net_config(None)                           # oracle/_ref/configs/network.yaml
            torch.manual_seed(0)
            enc, dec = get_network_params(False, H, W, input_channels=C, net_cfg=cfg)
            net = ED(False, enc, dec, 0.5, False, input_height=H, input_width=W).eval()
        finally:
            sys.path[:] = saved
        st = [torch.zeros(s) for s in state_shapes(H, W)]

        def one(i, st):
            out = net(xs[i % 2][None, None], *st)
            return list(out[1:])
    else:
        kind = "port"
        from oracle import torch_port as TP
        from src.lib.model.networks.model import ED  # noqa: F401  (parameter tree only; CPU tensors, never run)
        from src.lib.model.networks.net_params import get_network_params
        torch.manual_seed(0)
        enc, dec = get_network_params(False, H, W, input_channels=C)
        net = ED(False, enc, dec, 0.5, False, input_height=H, input_width=W)
        p = {k: v.detach() for k, v in net.state_dict().items()}
        st = [torch.zeros(s) for s in state_shapes(H, W)]

        def one(i, st):
            return TP.ed_step(p, xs[i % 2][None], st)[2]
    with torch.no_grad():
        for i in range(warmup):
            st = one(i, st)
        t0 = time.perf_counter()
        for i in range(steps):
            st = one(i, st)
        dt = time.perf_counter() - t0
    return H * W * steps / dt, dt / steps, torch.get_num_threads(), kind


def run_reference_arm(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    H, W, hist = a.height, a.width, a.hist
    steps = max(1, min(a.steps, 8))       # bounded sample: ~0.3 s per 500x500 step on the box's host cores
    warmup = max(1, min(a.warmup, 2))
    v, sec, cores, kind = cpu_reference_throughput(H, W, hist, steps, warmup)
    what = "unmodified reference modules (oracle/_ref)" if kind == "reference" else "op-for-op torch port of the reference"
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": a.gpus, "steps": steps,
            "warmup": warmup, "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": workload_name(H, W, 2 * hist + 3), "arm": f"CPU, {what}, torch {torch.__version__}, {cores} threads"},
            "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": kind,
                             "sample": f"{steps} ED steps at {H}x{W} after {warmup} warm-up, torch {torch.__version__} CPU"},
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    emit(line)


# ---------------------------------------------------------------------------------------------- our arm
def device_inputs(n, C, H, W, hist, dev, seed=42):
    """n distinct dense (C,H,W) inputs on the device: the notebook's synthetic event recipe (static maps + scalar rainfall
    history broadcast over the grid), built with torch so that chunks of it never pass through host memory."""
    g = torch.Generator(device="cpu").manual_seed(seed)
    maps = torch.rand((3, H, W), generator=g)
    maps[2] = (maps[2] > 0.95).float()
    rain = torch.rand((n + hist,), generator=g)
    cum = torch.cumsum(rain, 0) / 250.0 * 6.0
    x = torch.empty((n, C, H, W), device=dev)
    x[:, 2 * hist:] = maps.to(dev)
    for t in range(n):
        x[t, :hist] = rain[t:t + hist].to(dev)[:, None, None]
        x[t, hist:2 * hist] = cum[t:t + hist].to(dev)[:, None, None]
    return x


def run_ours(a):
    from urnn_b200 import _capi, ops
    from urnn_b200.runner import SequenceRunner
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py: no CUDA device; urnn_b200 has no CPU path (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist = None
    lib = _capi.load()
    if world > 1:
        # spatial sharding: row bands; the only per-step exchange is the in-kernel all-reduce of the normalisation
        # statistics over NVLink peer memory (no NCCL on the data path)
        import torch.distributed as dist
        from urnn_b200 import dist as ud
        dist.init_process_group("nccl", device_id=dev)
        ud.init_spatial_sharding()
    ops.set_default_math(a.math)
    hist = a.hist
    C = 2 * hist + 3
    gH, gW = a.height, a.width                      # the grid the metric is quoted on
    if world > 1 and a.scaling == "strong":
        if gH % (4 * world):
            raise SystemExit(f"--scaling strong: height {gH} must be a multiple of 4 * {world} (row bands aligned to the pools)")
        H, W, cells = gH // world, gW, gH * gW
    else:
        H, W, cells = gH, gW, world * gH * gW
    N = H * W
    net = build_net(H, W, C, a.math, dev)
    runner = SequenceRunner(net, H, W, C, math=a.math, use_graph=False)
    chunk = max(1, min(a.steps, 30 if N <= 512 * 512 else (8 if N <= 2048 * 2048 else 5)))
    xs_dev = device_inputs(chunk, C, H, W, hist, dev)

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
            torch.cuda.synchronize()

    def max_over_ranks(sec):
        if dist is None:
            return sec
        t = torch.tensor([sec], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # (1) device-resident inputs: K steps through urnn_ed_sequence_dev (states stay in the library's internal layout
    #     between the steps of a call; the fp32 NCHW states are converted at the two ends of every call)
    def run_steps(k, st):
        done = 0
        while done < k:
            n = min(chunk, k - done)
            _, _, st = runner.run_dev(xs_dev[:n], states=st, want_prob=False)
            done += n
        return st

    trace("built; warm-up")
    st = run_steps(max(a.warmup, 1), None)
    barrier()
    trace("warm-up done; timed steps")
    n0 = lib.urnn_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local, enabled=(rank == 0)) as clk:
        e0.record()
        st = run_steps(a.steps, st)
        e1.record()
        barrier()
        launches = lib.urnn_launch_count() - n0
        if a.steps * 1.1e-3 < 0.3:                  # keep the sampler alive long enough to see the clocks under load
            run_steps(max(1, int(0.3 / 1.1e-3) // chunk) * chunk, st)
            torch.cuda.synchronize()
    sec = max_over_ranks(e0.elapsed_time(e1) * 1e-3)
    value = cells * a.steps / sec
    trace(f"value leg done: {sec / a.steps * 1e3:.3f} ms/step")

    # sharded runs: the bands of a small grid against the same grid on one GPU (rank 0 recomputes it unsharded)
    sharded_parity = None
    if world > 1:
        sharded_parity = sharded_check(a, world, rank, dev, dist, hist)

    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": a.steps, "warmup": a.warmup,
            "ms_per_step": sec / a.steps * 1e3, "higher_is_better": True, "scaling": a.scaling if world > 1 else "weak", "vs_baseline": None,
            "dtype": {"fp32": "f32", "bf16": "bf16", "f16x3": "f16x3"}[a.math], "data": "synthetic",
            "config": {"workload": workload_name(gH, gW, C, world, a.scaling),
                       "l2": f"per-step working set (states in+out, dense input, LN affine) {(C + 2 * 188 + 160 + 1) * 4 * N / 1e6:.0f} MB per GPU > 126 MB L2; "
                             f"the dense inputs cycle through {chunk} distinct device tensors",
                       "weights": "random init, torch.manual_seed(0)", "math": a.math,
                       "parity": {"f16x3": "fp16 hi+lo split operands (22 bits), fp32 accumulate/statistics/pre-norm maps: atol 2e-5 vs the reference's goldens; "
                                           "500x500 T=180 vs the fp32 path: rms 3e-5, max |d state| 0.014 (tests/test_gpu_x3.py)",
                                  "fp32": "fp32 FFMA: atol 1e-5 / rtol 1e-4 vs the reference",
                                  "bf16": "single-pass bf16 operands: short horizons only (outside the config-3 tolerance at T=180)"}[a.math],
                       "sharding": "row bands, in-kernel NVLink statistic all-reduce" if world > 1 else "none",
                       "value_path": f"urnn_ed_sequence_dev (C ABI, device buffers), {chunk}-step calls"
                                     + ("; encoder(t+1) overlaps decoder+head(t) on two streams" if world == 1 and a.math == "f16x3" and H * W <= (4 << 20) else "")},
            "gpu_launches": int(launches), "clocks": clk.summary()}
    if sharded_parity is not None:
        line["sharded_parity"] = sharded_parity

    if a.value_only:
        if rank == 0:
            emit(line)
        shutdown(dist)
        return

    # (2) end to end, the reference's workflow (test.py:447 uploads the event once; test.py:356-375 loops and reads every
    #     step's depth map back): urnn_ed_event_host -- raw maps + scalar rainfall from pinned HOST memory once per event,
    #     per-step input assembly folded into the stage-1 stem, one D2H copy of the (H,W) depth map per step
    rng = np.random.RandomState(42)
    ev_maps = [torch.from_numpy(m.astype(np.float32)).pin_memory() for m in (rng.rand(H, W) * 10.0, rng.rand(H, W), (rng.rand(H, W) > 0.95) * 1.0)]
    rain = torch.from_numpy((rng.rand(a.steps) * 6.0).astype(np.float32))
    ev_out = torch.empty((a.steps, H, W), dtype=torch.float32).pin_memory()
    runner.run_event_host(*ev_maps, rain[:2], torch.cumsum(rain[:2], 0), hist, 6.0, 250.0, out_host=ev_out[:2])
    barrier()
    t0 = time.perf_counter()
    runner.run_event_host(*ev_maps, rain, torch.cumsum(rain, 0), hist, 6.0, 250.0, out_host=ev_out)
    torch.cuda.synchronize()
    sec_ev = max_over_ranks(time.perf_counter() - t0)
    e2e_event = cells * a.steps / sec_ev
    trace("event leg done")

    # (2b) dense inputs from the host every step (spatial-rainfall datasets): urnn_ed_sequence_host
    hchunk = min(a.steps, 16)
    in_host = torch.empty((hchunk, C, H, W), dtype=torch.float32).pin_memory()
    in_host.copy_(xs_dev[:1].expand(hchunk, -1, -1, -1) if chunk < hchunk else xs_dev[:hchunk])
    out_host = torch.empty((hchunk, H, W), dtype=torch.float32).pin_memory()
    runner.run_host(in_host[:2], out_host[:2])
    barrier()
    t0 = time.perf_counter()
    done, sth = 0, None
    while done < a.steps:
        n = min(hchunk, a.steps - done)
        _, sth = runner.run_host(in_host[:n], out_host[:n], states=sth)
        done += n
    torch.cuda.synchronize()
    sec_dense = max_over_ranks(time.perf_counter() - t0)
    e2e_dense = cells * a.steps / sec_dense
    trace("dense e2e leg done")

    # (3) roofline: per-launch CUDA-event times inside the running step (urnn_ed_profile_dev).  Algorithmic bytes per cell
    #     step = (C_x + C_h + F) * 4: decoder stage 1 (in 96 + e 64 + d 64 -> 64) and encoder stage 1 (16 + 64 -> 64)
    peak, peak_src = measured_peaks()
    roofline = None
    if a.math == "f16x3":
        tab = runner.profile_dev(xs_dev[:min(chunk, 12)])
        ops_ms = dict(tab)

        def cell(prefix, elems):
            ms = sum(v for k, v in tab if k.startswith(prefix + "."))
            alg = elems * 4 * N
            return {"ms_per_launch_group": ms, "launches": [k for k, _ in tab if k.startswith(prefix + ".")], "algorithmic_bytes": alg,
                    "achieved": alg / (ms * 1e-3) / 1e9, "frac": alg / (ms * 1e-3) / 1e9 / peak}
        dec1, enc1 = cell("dec1", 96 + 64 + 64 + 64), cell("enc1", 16 + 64 + 64)
        traffic, tpath = None, os.path.join(ROOT, "profiles", "r2_traffic_cells.json")
        if os.path.exists(tpath) and (H, W) == (H_DEF, W_DEF):
            with open(tpath) as f:
                traffic = json.load(f)
        step_alg = (C + 2 * 188 + 160 + 1) * 4 * N
        roofline = {"bound": "hbm", "kernel": f"decoder stage-1 Skip-ConvGRU cell step (sweep A, sweep B, blend) at {H}x{W}, math={a.math}",
                    "achieved": dec1["achieved"], "peak": peak, "unit": "GB/s", "frac": dec1["frac"],
                    "traffic": traffic["dec1"] if traffic else None, "traffic_note": traffic.get("note") if traffic else None,
                    "peak_source": peak_src,
                    "algorithmic_bytes_per_launch": dec1["algorithmic_bytes"], "ms_per_launch": dec1["ms_per_launch_group"],
                    "worst_cell": dict(enc1, kernel="encoder stage-1 ConvGRU cell step", traffic=traffic["enc1"] if traffic else None),
                    "per_launch_us": {k: round(v * 1e3, 1) for k, v in tab},
                    "whole_step": {"algorithmic_bytes": step_alg, "achieved_GBps": step_alg * a.steps / sec / 1e9,
                                   "frac": step_alg * a.steps / sec / 1e9 / peak}}
    trace("roofline leg done")

    # (4) CPU baseline beside it (rank 0, N = 1 only: at N > 1 the other ranks would be spinning in a barrier on the same cores)
    cpu = None
    if rank == 0 and world == 1 and not a.no_cpu_baseline:
        v, s_, cores, kind = cpu_reference_throughput(H, W, hist, 6, 2)
        cpu = {"value": v, "unit": UNIT, "cores": cores, "kind": kind,
               "sample": f"6 ED steps at {H}x{W} after 2 warm-up ({s_ * 1e3:.0f} ms/step), torch {torch.__version__} CPU"}

    if rank == 0:
        line["e2e"] = {"value": e2e_event, "unit": UNIT, "h2d_bytes_per_step": (3 * N * 4 + 2 * a.steps * 4) / a.steps, "d2h_bytes_per_step": N * 4,
                       "ms_per_step": sec_ev / a.steps * 1e3, "h2d_bytes_per_event": 3 * N * 4 + 2 * a.steps * 4,
                       "path": "urnn_ed_event_host (C ABI, host buffers): the event is uploaded once (test.py:447), every step's depth map is copied back"}
        line["e2e_dense"] = {"value": e2e_dense, "unit": UNIT, "h2d_bytes_per_step": C * N * 4, "d2h_bytes_per_step": N * 4,
                             "ms_per_step": sec_dense / a.steps * 1e3, "path": f"urnn_ed_sequence_host, {hchunk}-step calls"}
        line["roofline"] = roofline
        line["cpu_baseline"] = cpu
        emit(line)
    shutdown(dist)


def run_train(a):
    """BASELINE config 4: SWP training windows at the location1 grid.  One "step" = one window: seq_num time steps forward
    through the drop-in modules with autograd (main.py:674-684), loss, backward (recompute-in-backward CUDA kernels, fp32),
    all-reduce of the replicated weight gradients + global clip norm when sharded (main.py:756-761), Adam step."""
    from urnn_b200 import _capi, ops
    world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0")); local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        from urnn_b200 import dist as ud
        dist.init_process_group("nccl", device_id=dev)
        ud.init_spatial_sharding()
    from urnn_b200 import dist as ud
    ops.set_default_math("fp32")
    hist = a.hist; C = 2 * hist + 3; H, W = a.height, a.width; S = a.seq_num
    net = build_net(H, W, C, "fp32", dev).train()
    opt = torch.optim.Adam(net.parameters(), lr=0.01)
    xs = device_inputs(S, C, H, W, hist, dev)
    torch.manual_seed(7 + rank)
    label = torch.rand(S, H, W, device=dev) * 0.3
    label[label < 0.25] = 0
    lib = _capi.load()

    def window():
        opt.zero_grad(set_to_none=True)
        st = [torch.zeros(s, device=dev) for s in state_shapes(H, W)]
        regs = []
        for t in range(S):
            out, *st = net(xs[t][None, None], *st)
            regs.append(out)
        reg = torch.cat(regs, dim=1)[0]
        wet = (label > 0).float()
        loss = (((reg - label) ** 2) * (1.0 + 19.0 * wet)).sum() / (S * H * W * world)      # WMSE-style weighting, global normalisation
        loss.backward()
        if world > 1:
            ud.allreduce_window_gradients(net)
            ud.clip_grad_norm_sharded(net, 1.0)
        else:
            torch.nn.utils.clip_grad_norm_(net.parameters(), 1.0)
        opt.step()
        return loss

    for _ in range(max(1, a.warmup)):
        window()
    torch.cuda.synchronize()
    if dist is not None:
        dist.barrier(); torch.cuda.synchronize()
    n0 = lib.urnn_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local, enabled=(rank == 0)) as clk:
        e0.record()
        for _ in range(a.steps):
            loss = window()
        e1.record()
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier(); torch.cuda.synchronize()
    sec = e0.elapsed_time(e1) * 1e-3
    if dist is not None:
        t = torch.tensor([sec], device=dev, dtype=torch.float64); dist.all_reduce(t, op=dist.ReduceOp.MAX); sec = float(t.item())
    if rank == 0:
        emit({"metric": "grid-cells*steps/s (SWP window: fwd + bwd + grad all-reduce + clip + Adam)", "mode": "train",
              "value": world * H * W * S * a.steps / sec, "unit": UNIT, "n_gpus": world, "steps": a.steps, "warmup": a.warmup,
              "ms_per_step": sec / a.steps * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
              "config": {"workload": f"location1 SWP training window: {world * H}x{W}" + (f" ({world} row bands)" if world > 1 else "") +
                                     f", C_in={C}, seq_num={S}, Adam lr 0.01, grad_clip 1.0, weighted-MSE loss",
                         "sharding": "row bands; in-kernel NVLink statistic exchange (fwd + bwd), NCCL all-reduce of 420146 replicated gradients per window" if world > 1 else "none",
                         "published": "README.md:626 (RTX 4090, derived): ~1.7 M cells*steps/s fwd+bwd"},
              "gpu_launches": int(lib.urnn_launch_count() - n0), "clocks": clk.summary(), "loss": float(loss)})
    shutdown(dist)


def shutdown(dist):
    if dist is not None:
        from urnn_b200 import dist as ud
        ud.shutdown_spatial_sharding()
        dist.destroy_process_group()


def sharded_check(a, world, rank, dev, dist, hist):
    """Multi-GPU correctness made visible in the bench line: a small grid split into `world` row bands (through this run's
    communicator) against the same grid computed unsharded on every rank, same weights and inputs."""
    from urnn_b200 import dist as ud
    from urnn_b200.runner import SequenceRunner
    C = 2 * hist + 3
    bh, W, T = 32, 64, 3                      # band height 32 -> a (32*world) x 64 grid
    H = bh * world
    torch.manual_seed(0)
    full = build_net(H, W, C, a.math, dev)    # same seed on every rank: identical replicated weights
    xs = device_inputs(T, C, H, W, hist, dev, seed=7)
    sd = ud.shard_state_dict(full.state_dict(), world, rank)
    band = build_net(bh, W, C, a.math, dev)
    band.load_state_dict(sd)
    r0, r1 = rank * bh, (rank + 1) * bh
    depth_b, _, _ = SequenceRunner(band, bh, W, C).run_dev(xs[:, :, r0:r1].contiguous(), want_prob=False)
    torch.cuda.synchronize()
    dist.barrier()
    # unsharded run: the communicator is switched off for it
    ud.shutdown_spatial_sharding()
    depth_f, _, _ = SequenceRunner(full, H, W, C).run_dev(xs, want_prob=False)
    torch.cuda.synchronize()
    err = (depth_b - depth_f[:, r0:r1]).abs().max().reshape(1).double()
    dist.all_reduce(err, op=dist.ReduceOp.MAX)
    ud.init_spatial_sharding()
    return {"grid": f"{H}x{W} in {world} bands, T={T}", "max_abs_err_vs_unsharded": float(err.item())}


class StdoutToStderr:
    """stdout must carry exactly one JSON line, but NCCL prints its banner ("NCCL version ...") to fd 1 when the box
    exports NCCL_DEBUG: while the benchmark runs, fd 1 points at stderr; `restore()` is called right before the line."""

    def __init__(self):
        sys.stdout.flush()
        self.saved = os.dup(1)
        os.dup2(2, 1)

    def restore(self):
        if self.saved is not None:
            sys.stdout.flush()
            os.dup2(self.saved, 1)
            os.close(self.saved)
            self.saved = None


_REDIRECT = None


def emit(line):
    if _REDIRECT is not None:
        _REDIRECT.restore()
    print(json.dumps(line), flush=True)


def main():
    global _REDIRECT
    a = parse()
    # a wedged GPU must not hang the caller: dump the Python stack and exit instead
    import faulthandler
    faulthandler.dump_traceback_later(a.watchdog, exit=True)
    if int(os.environ.get("WORLD_SIZE", "1")) > 1:
        _REDIRECT = StdoutToStderr()
    if a.impl == "reference":
        run_reference_arm(a)
    elif a.mode == "train":
        run_train(a)
    else:
        run_ours(a)


if __name__ == "__main__":
    main()
