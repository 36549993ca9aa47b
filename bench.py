#!/usr/bin/env python
"""
bench.py -- grid-cells*steps/s of the U-RNN encoder-decoder time step (flood-depth forward) on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--math bf16|fp32]

One "step" = one ED time step (reference model.py:65-121) over the whole H x W grid.  Workload at N=1: the
location1 full-resolution grid (500 x 500, 2 m / 1 min, historical_nums=30 => C_in=63), synthetic event,
seeded random-init weights (no dataset / checkpoint offline).  Prints ONE JSON line.

  value     : cells*steps/s with the per-step inputs already resident in HBM (ring of distinct inputs)
  e2e       : same metric through the C ABI with HOST buffers (urnn_ed_sequence_host): per step one H2D copy of the
              (C_in,H,W) input from pinned memory and one D2H copy of the (H,W) depth map, inside the timed region
  roofline  : dominant op timed alone with CUDA events (algorithmic bytes / time vs MEASURED_PEAKS.json hbm_gbs)
  cpu_baseline : the functional-torch port of the reference (oracle/torch_port.py) on this box's host cores
  --impl reference : that same CPU port as its own arm (the Python reference cannot travel to the GPU box)
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
    ap.add_argument("--math", default=os.environ.get("URNN_MATH", "bf16"), choices=["fp32", "bf16"])
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
    """nvidia-smi clocks / throttle reasons sampled every 50 ms while the timed region runs."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.index, self.proc, self.lines = index, None, []

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "50"],
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
        if self.proc is not None:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()

    def summary(self):
        sm, mx, reasons = [], 0, set()
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
                "reasons": sorted(reasons), "samples": len(sm)}


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
def cpu_port_throughput(H, W, hist, steps, warmup):
    """The functional-torch CPU port of the reference's step, all host threads."""
    from oracle import torch_port as TP
    from src.lib.model.networks.model import ED  # noqa: F401  (parameter tree only; CPU tensors, never run)
    from src.lib.model.networks.net_params import get_network_params
    C = 2 * hist + 3
    torch.manual_seed(0)
    enc, dec = get_network_params(False, H, W, input_channels=C)
    net = ED(False, enc, dec, 0.5, False, input_height=H, input_width=W)
    p = {k: v.detach() for k, v in net.state_dict().items()}
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    xs = torch.from_numpy(synthetic_inputs(H, W, hist, 2))
    st = [torch.zeros(s) for s in state_shapes(H, W)]
    with torch.no_grad():
        for i in range(warmup):
            _, _, st = TP.ed_step(p, xs[i % 2][None], st)
        t0 = time.perf_counter()
        for i in range(steps):
            _, _, st = TP.ed_step(p, xs[i % 2][None], st)
        dt = time.perf_counter() - t0
    return H * W * steps / dt, dt / steps, torch.get_num_threads()


def run_reference_arm(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    H, W, hist = a.height, a.width, a.hist
    steps = max(1, min(a.steps, 8))       # bounded sample: ~1.5 s per 500x500 step on 8 cores
    warmup = max(1, min(a.warmup, 2))
    v, sec, cores = cpu_port_throughput(H, W, hist, steps, warmup)
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": a.gpus, "steps": steps,
            "warmup": warmup, "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"location1 full-res {H}x{W}, C_in={2 * hist + 3}, ED step forward, CPU torch port of the reference"},
            "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
                             "sample": f"{steps} ED steps at {H}x{W} after {warmup} warm-up, torch {torch.__version__} CPU"},
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    emit(line)


# ---------------------------------------------------------------------------------------------- our arm
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
        # spatial sharding: every rank owns one H x W band of a (world*H) x W grid; the only per-step exchange is the
        # in-kernel all-reduce of the normalisation statistics over NVLink peer memory (no NCCL on the data path)
        import torch.distributed as dist
        from urnn_b200 import dist as ud
        dist.init_process_group("nccl", device_id=dev)
        ud.init_spatial_sharding()
    ops.set_default_math(a.math)
    H, W, hist = a.height, a.width, a.hist
    C = 2 * hist + 3
    N = H * W
    net = build_net(H, W, C, a.math, dev)
    runner = SequenceRunner(net, H, W, C, math=a.math, use_graph=False)
    desc, params = runner.desc, net.ed_params()
    ring = 2 if a.value_only else 8
    xs_host = torch.from_numpy(synthetic_inputs(H, W, hist, ring)).pin_memory()
    xs_dev = xs_host.to(dev)
    out = torch.empty((2, H, W), device=dev)

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

    # (1) device-resident inputs: K steps of the whole encoder-decoder step, states ping-pong in HBM
    def step(i):
        ops.ed_step_fwd(desc, params, xs_dev[i % ring], runner.states[i & 1], runner.states[(i & 1) ^ 1], out, runner.ws)

    trace("built; warm-up")
    with torch.no_grad():
        for i in range(a.warmup):
            step(i)
        barrier()
        trace("warm-up done; timed steps")
        n0 = lib.urnn_launch_count()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with ClockSampler(local) as clk:
            e0.record()
            for i in range(a.steps):
                step(i)
            e1.record()
            barrier()
        launches = lib.urnn_launch_count() - n0
    sec = max_over_ranks(e0.elapsed_time(e1) * 1e-3)
    value = world * N * a.steps / sec
    trace(f"value leg done: {sec / a.steps * 1e3:.3f} ms/step")

    if a.value_only:
        if rank == 0:
            emit({"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": a.steps, "warmup": a.warmup,
                              "ms_per_step": sec / a.steps * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                              "dtype": {"fp32": "f32", "bf16": "bf16"}[a.math], "data": "synthetic",
                              "config": {"workload": f"{world * H}x{W}, C_in={C}, ED step forward, value only", "math": a.math},
                              "gpu_launches": int(launches), "clocks": clk.summary()})
        if dist is not None:
            from urnn_b200 import dist as ud
            ud.shutdown_spatial_sharding()
            dist.destroy_process_group()
        return

    # (2) end to end through the C ABI with HOST buffers (urnn_ed_sequence_host): per step one H2D copy of the
    #     (C_in,H,W) input from pinned memory and one D2H copy of the (H,W) depth map, overlapped with compute
    chunk = min(a.steps, 16)
    in_host = torch.empty((chunk, C, H, W), dtype=torch.float32).pin_memory()
    for i in range(chunk):
        in_host[i].copy_(xs_host[i % ring])
    out_host = torch.empty((chunk, H, W), dtype=torch.float32).pin_memory()
    runner.run_host(in_host[:max(1, min(a.warmup, chunk))], out_host[:max(1, min(a.warmup, chunk))])
    barrier()
    t0 = time.perf_counter()
    done = 0
    while done < a.steps:
        n = min(chunk, a.steps - done)
        runner.run_host(in_host[:n], out_host[:n], states=runner.states[0])
        done += n
    torch.cuda.synchronize()
    sec_e2e = max_over_ranks(time.perf_counter() - t0)
    e2e = world * N * a.steps / sec_e2e
    trace("e2e leg done")

    # (2b) the reference's own workflow end to end (test.py:447-375): the event dict comes from the HOST once (3 static
    #      maps + the scalar rainfall series), inputs are assembled on the device every step (here: folded into the
    #      stage-1 bias), depth maps return to the host every step.  urnn_ed_event_host, row f-1 of SURVEY.md 8.
    rng = np.random.RandomState(42)
    ev_maps = [torch.from_numpy(a.astype(np.float32)).pin_memory() for a in (rng.rand(H, W) * 10.0, rng.rand(H, W), (rng.rand(H, W) > 0.95) * 1.0)]
    rain = torch.from_numpy((rng.rand(a.steps) * 6.0).astype(np.float32))
    ev_out = torch.empty((a.steps, H, W), dtype=torch.float32).pin_memory()
    runner.run_event_host(*ev_maps, rain[:2], torch.cumsum(rain[:2], 0), hist, 6.0, 250.0, out_host=ev_out[:2])
    barrier()
    t0 = time.perf_counter()
    runner.run_event_host(*ev_maps, rain, torch.cumsum(rain, 0), hist, 6.0, 250.0, out_host=ev_out)
    torch.cuda.synchronize()
    sec_ev = max_over_ranks(time.perf_counter() - t0)
    e2e_event = world * N * a.steps / sec_ev
    trace("event leg done")

    # (3) roofline of the dominant op: the full-resolution decoder Skip-ConvGRU cell (36 % of the step's FLOPs and
    #     its largest kernels); algorithmic bytes = (C_x + C_e + C_d + F) * 4 per cell
    cell = net.decoder.rnn1
    F, Cx = cell.num_features, cell.input_channels
    x = torch.rand(Cx, H, W, device=dev); e = torch.rand(F, H, W, device=dev)
    hs = [torch.rand(F, H, W, device=dev) for _ in range(4)]      # rotate buffers: > L2 together with the workspace
    with torch.no_grad():
        for i in range(3):
            cell.step(x, e, hs[i % 4])
        torch.cuda.synchronize()
        c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 20
        c0.record()
        for i in range(reps):
            cell.step(x, e, hs[i % 4])
        c1.record()
        torch.cuda.synchronize()
    cell_sec = c0.elapsed_time(c1) * 1e-3 / reps
    trace("roofline leg done")
    alg_bytes = (Cx + 2 * F + F) * 4 * N
    peak, peak_src = measured_peaks()
    achieved = alg_bytes / cell_sec / 1e9
    # measured DRAM traffic of the same three launches: one `ncu --set full` capture (tools/prof_cell.py), committed
    traffic, tpath = None, os.path.join(ROOT, "profiles", "r1_traffic_dec1_cell.json")
    if os.path.exists(tpath) and (H, W) == (H_DEF, W_DEF) and a.math == "bf16":
        with open(tpath) as f:
            traffic = float(json.load(f)["dram_bytes_per_cell_step"])
    roofline = {"bound": "hbm", "kernel": f"decoder stage-1 Skip-ConvGRU cell step (in={Cx}, F={F}; 3 sweeps) at {H}x{W}, math={a.math}",
                "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                "peak_source": peak_src, "algorithmic_bytes_per_launch": alg_bytes, "ms_per_launch": cell_sec * 1e3,
                "whole_step": {"algorithmic_bytes": (C + 2 * 188 + 160 + 1) * 4 * N,
                               "achieved_GBps": (C + 2 * 188 + 160 + 1) * 4 * N * world * a.steps / sec / 1e9 / world,
                               "frac": (C + 2 * 188 + 160 + 1) * 4 * N * a.steps / sec / 1e9 / peak}}

    # (4) CPU baseline beside it
    cpu = None
    if rank == 0 and not a.no_cpu_baseline:
        v, s_, cores = cpu_port_throughput(H, W, hist, 6, 2)
        cpu = {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
               "sample": f"6 ED steps at {H}x{W} after 2 warm-up ({s_ * 1e3:.0f} ms/step), torch {torch.__version__} CPU"}

    if rank == 0:
        grid = f"{world * H}x{W} ({world} row bands of {H}x{W})" if world > 1 else f"{H}x{W}"
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": a.steps, "warmup": a.warmup,
                "ms_per_step": sec / a.steps * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": {"fp32": "f32", "bf16": "bf16"}[a.math], "data": "synthetic",
                "config": {"workload": f"location1 full-res {grid}, C_in={C}, ED step forward (6 ConvGRU cells + stems + head)",
                           "l2": "per-step working set ~600 MB (states in+out, inputs, LN affine) > 126 MB L2; inputs cycle through a ring of 8",
                           "weights": "random init, torch.manual_seed(0)", "math": a.math,
                           "parity": "per-step parity tests in tests/ (fp32: 1e-5 vs the reference; bf16: the mode's model); bf16 vs fp32 drift over T=180 "
                                     "with random-init weights: profiles/r1_config3_drift.json (R2 0.973)" if a.math == "bf16" else "fp32: atol 1e-5 / rtol 1e-4 vs the reference",

                           "sharding": "row bands, in-kernel NVLink statistic all-reduce" if world > 1 else "none",
                           "e2e_path": f"urnn_ed_sequence_host (C ABI, host buffers), {chunk}-step calls"},
                "e2e": {"value": e2e, "unit": UNIT, "h2d_bytes_per_step": C * N * 4, "d2h_bytes_per_step": N * 4,
                        "ms_per_step": sec_e2e / a.steps * 1e3},
                "e2e_event": {"value": e2e_event, "unit": UNIT, "h2d_bytes_per_event": 3 * N * 4 + 2 * a.steps * 4,
                              "d2h_bytes_per_step": N * 4, "ms_per_step": sec_ev / a.steps * 1e3,
                              "path": "urnn_ed_event_host: raw maps + scalar rainfall from the host once, per-step input assembly fused into the stage-1 stem"},
                "gpu_launches": int(launches), "clocks": clk.summary(), "roofline": roofline, "cpu_baseline": cpu}
        emit(line)
    if dist is not None:
        from urnn_b200 import dist as ud
        ud.shutdown_spatial_sharding()
        dist.destroy_process_group()


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
    else:
        run_ours(a)


if __name__ == "__main__":
    main()
