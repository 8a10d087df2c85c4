#!/usr/bin/env python
"""bench.py -- path samples/s of the B200 volume path tracer on BASELINE.json's metric config.

    python bench.py --gpus N --steps K --warmup W [--impl reference]

Workload at every N: BASELINE.json configs[1] -- data/smoke.brick + data/lut.txt transfer function
(pathtracer_brick_tf), 1920x1080, 128 bounces. One "step" = one batch of --spp-per-step samples for every
pixel (the full config is 1024 spp = 1024/spp-per-step such steps; samples are independent).
N > 1: spp-sliced weak scaling -- every rank traces its own --spp-per-step slice of sample indices for all
pixels into a SUM buffer, the float4 buffers are reduced to rank 0 with NCCL (the one real exchange step).

value  : whole-job samples/s with volume/env/LUT resident in HBM (CUDA events around the K steps, max over ranks)
e2e    : same metric through the C ABI with HOST buffers: per step the brick grid, environment and LUT are
         uploaded from pinned host memory, traced, and the RGBA32F image is read back (H2D/D2H inside the timed region);
         steps rotate over four contexts with asynchronous uploads so that read-backs, uploads and traces overlap
roofline: algorithmic bytes (event counters x per-event bytes) / kernel time vs the measured HBM peak
cpu_baseline: the CPU oracle (port of the reference shaders) on a bounded sample of the same workload
"""
from __future__ import annotations

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
sys.path.insert(0, os.path.join(ROOT, "tests"))

ASSETS = os.path.join(ROOT, "tests", "golden", "assets")
W, H, BOUNCES, FULL_SPP = 1920, 1080, 128, 1024
METRIC = "path samples/sec (1080p, 128 bounces)"
UNIT = "samples/s"


def load_workload(w=W, h=H):
    from volren_b200 import formats
    from helpers import default_scene
    grid = formats.load_brick(os.path.join(ASSETS, "smoke.brick"))
    env = formats.load_hdr(os.path.join(ASSETS, "table_mountain_2_puresky_1k.hdr"))
    lut = formats.lut_for_upload(formats.load_lut_txt(os.path.join(ASSETS, "lut.txt")))
    # `./volren data/smoke.brick <hdr> data/lut.txt -w 1920 -h 1080 --render --bounces 128`
    params = default_scene(grid, w, h, bounces=BOUNCES, use_tf=True)
    return grid, env, lut, params


# CPU legs: the SAME frame (camera, fov, aspect -> the same mix of empty and dense pixels) at 1/4 resolution per axis
CPU_W, CPU_H = W // 4, H // 4


def algorithmic_bytes(c: dict, use_tf: bool) -> float:
    """bytes the algorithm must touch, from the event counters (per launch)."""
    if use_tf:
        per_maj, per_dens = 4 + 32, 8 * 9 + 32
    else:
        per_maj, per_dens = 4, 9
    return (per_maj * c["n_maj"] + per_dens * c["n_dens"] + 9 * c["n_emis"] + 200 * c["n_nee"] + 100 * c["n_env"]
            + 32 * c["n_samples"])


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return json.load(open(p)), "measured"
    return {"hbm_gbs": 6650.0}, "fallback"


class ClockSampler:
    """SM clock and throttle reasons DURING the timed region, polled through NVML every few ms from a thread (the timed
    region of a short run is shorter than one `nvidia-smi -lms` period); falls back to nvidia-smi when NVML is missing."""
    REASONS = {"hw_slowdown": 0x8, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20, "sw_power_cap": 0x4}

    def __init__(self, index=0, period_s=0.002):
        self.index, self.period, self.sm, self.mask, self.max_mhz = index, period_s, [], 0, None
        self.stop_flag = threading.Event()
        self.thread = None
        try:
            import pynvml
            pynvml.nvmlInit()
            # CUDA_VISIBLE_DEVICES may remap indices: match by PCI bus id of the torch device when possible
            import torch
            try:
                bus = torch.cuda.get_device_properties(index).pci_bus_id
                dom = getattr(torch.cuda.get_device_properties(index), "pci_domain_id", 0)
                dev = getattr(torch.cuda.get_device_properties(index), "pci_device_id", 0)
                self.handle = pynvml.nvmlDeviceGetHandleByPciBusId(f"{dom:08x}:{bus:02x}:{dev:02x}.0".encode())
            except Exception:
                self.handle = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.nvml = pynvml
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.handle, pynvml.NVML_CLOCK_SM))
        except Exception:
            self.nvml = None

    def _poll(self):
        n = self.nvml
        while not self.stop_flag.is_set():
            try:
                self.sm.append(float(n.nvmlDeviceGetClockInfo(self.handle, n.NVML_CLOCK_SM)))
                self.mask |= int(n.nvmlDeviceGetCurrentClocksEventReasons(self.handle))
            except Exception:
                try:
                    self.mask |= int(n.nvmlDeviceGetCurrentClocksThrottleReasons(self.handle))
                except Exception:
                    pass
            time.sleep(self.period)

    def start(self):
        if self.nvml is not None:
            self.thread = threading.Thread(target=self._poll, daemon=True)
            self.thread.start()

    def stop(self):
        if self.nvml is None:
            try:
                q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
                f = [x.strip() for x in subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q, "--format=csv,noheader,nounits"],
                                                       capture_output=True, text=True, timeout=10).stdout.strip().split(",")]
                names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
                return {"sm_mhz": float(f[0]), "sm_max_mhz": float(f[1]), "reasons": [n for n, v in zip(names, f[2:6]) if v.lower().startswith("active")],
                        "samples": 1, "how": "nvidia-smi once after the timed region (NVML unavailable)"}
            except Exception:
                return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"], "samples": 0}
        self.stop_flag.set()
        self.thread.join(timeout=1)
        return {"sm_mhz": float(np.median(self.sm)) if self.sm else None, "sm_max_mhz": self.max_mhz,
                "reasons": sorted(k for k, bit in self.REASONS.items() if self.mask & bit), "samples": len(self.sm), "how": "NVML poll every 2 ms during the timed region"}


def run_reference(args, rank, world):
    """--impl reference: the reference's CPU implementation of the path. The GLSL renderer needs a GL context that
    neither this container nor the GPU box has, so this is the oracle port of the shaders with all host threads
    (kind = "port"); each step is a bounded crop of the same workload."""
    if rank != 0:
        return
    from oracle.binding import Oracle
    grid, env, lut, params = load_workload(CPU_W, CPU_H)
    o = Oracle()
    pyr = o.env_build(env)
    sc = o.make_scene(grid, env, pyr, lut=lut)
    cores = len(os.sched_getaffinity(0))     # all host threads, whatever OMP_NUM_THREADS torchrun exported
    # bounded sample: the whole frame at 480x270, spp per step sized for ~3 s per step
    color = np.zeros((CPU_H, CPU_W, 4), np.float32)
    o.trace(sc, params, 1, 1, color=color, n_threads=cores)
    t0 = time.perf_counter()
    o.trace(sc, params, 1, 4, color=color, n_threads=cores)
    probe = (time.perf_counter() - t0) / 4
    spp = max(1, int(min(3.0, 90.0 / max(args.steps, 1)) / max(probe, 1e-4)))      # ~3 s per step, the whole run <= ~90 s
    for i in range(args.warmup):
        o.trace(sc, params, 1 + i, 1, color=color, n_threads=cores)
    t0 = time.perf_counter()
    for k in range(args.steps):
        o.trace(sc, params, 1 + k * spp, spp, color=color, n_threads=cores)
    dt = time.perf_counter() - t0
    samples = CPU_W * CPU_H * spp * args.steps
    v = samples / dt
    sample = f"the 1080p frame rendered at {CPU_W}x{CPU_H} (same camera), {spp} spp per step, {args.steps} steps"
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "reference assets (smoke.brick, lut.txt, hdr) committed under tests/golden/assets",
        "config": {"workload": "configs[1]: smoke.brick + lut.txt TF (pathtracer_brick_tf), 1920x1080, 128 bounces", "sample": sample},
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def cpu_baseline_worker():
    """cpu_baseline leg: the oracle (port of the reference shaders) on a bounded sample of the bench workload."""
    from oracle.binding import Oracle
    grid, env, lut, params = load_workload(CPU_W, CPU_H)
    o = Oracle()
    pyr = o.env_build(env)
    sc = o.make_scene(grid, env, pyr, lut=lut)
    img = np.zeros((CPU_H, CPU_W, 4), np.float32)
    cores = len(os.sched_getaffinity(0))
    o.trace(sc, params, 1, 1, color=img, n_threads=cores)
    t0 = time.perf_counter()
    o.trace(sc, params, 1, 4, color=img, n_threads=cores)
    probe = (time.perf_counter() - t0) / 4
    spp = max(1, int(12.0 / max(probe, 1e-4)))
    t0 = time.perf_counter()
    o.trace(sc, params, 2, spp, color=img, n_threads=cores)
    dt = time.perf_counter() - t0
    print(json.dumps({"value": CPU_W * CPU_H * spp / dt, "unit": UNIT, "cores": cores, "kind": "port",
                      "sample": f"the 1080p frame rendered at {CPU_W}x{CPU_H} (same camera), {spp} spp (oracle/vr_oracle.c, OpenMP)"}))


def main():
    if "--cpu-baseline-worker" in sys.argv:
        cpu_baseline_worker()
        return
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--spp-per-step", type=int, default=32)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    import torch.distributed as dist
    import volren_b200 as vr

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: volren_b200 has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    dev = torch.device("cuda", local_rank)
    grid, env, lut, params = load_workload()
    S = args.spp_per_step

    ctx = vr.Context(local_rank)
    stream = torch.cuda.current_stream()
    ctx.set_stream(stream.cuda_stream)          # kernels run on torch's current stream: torch.cuda.Event sees them
    ctx.resize(W, H)
    color = torch.zeros((H, W, 4), dtype=torch.float32, device=dev)
    ctx.bind_color(color.data_ptr())            # NCCL reduces this tensor in place
    ctx.grid_upload_brick(grid)
    ctx.env_upload(env)
    ctx.tf_upload(lut)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)   # > 126 MB L2

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    from volren_b200.multigpu import PartitionedRenderer
    kev_cur = [None]

    def trace_fn(first, n, tile, accum):
        if kev_cur[0] is not None:
            kev_cur[0][0].record()
        ctx.trace(params, first, n, tile=tile, accum_mode=accum)
        if kev_cur[0] is not None:
            kev_cur[0][1].record()

    # spp slices: rank r traces S of the S*world samples of a step into a SUM buffer, one NCCL reduce(SUM) to rank 0
    pr = PartitionedRenderer(color, trace_fn, partition="spp")

    def step(last=True):
        """one step: S more samples per pixel on every rank (weak scaling). The float4 accumulation buffers are reduced to
        rank 0 ONCE per frame: by the last step of the frame, inside the timed region."""
        pr.render(S * world, reduce=last)

    # ---- counting pass (defines the algorithmic bytes of one launch) ----
    ctx.set_counting(True)
    ctx.trace(params, 1, S)
    counters = ctx.get_counters().as_dict()
    alg_bytes = algorithmic_bytes(counters, use_tf=True)
    # the same with the production launch's exact culling kept (hidden environment: pixels / tiles that cannot produce a
    # non-zero sample are not traced): the events -- and bytes -- the timed kernel really executes
    ctx.set_option("count_culled", 1)
    ctx.set_counting(True)                      # resets the counters
    ctx.clear()
    ctx.trace(params, 1, S)
    counters_exec = ctx.get_counters().as_dict()
    ctx.set_option("count_culled", 0)
    ctx.set_counting(False)
    exec_bytes = algorithmic_bytes(counters_exec, use_tf=True) + 32 * (counters["n_samples"] - counters_exec["n_samples"])   # folded zeros still cost the RMW

    # ---- device-resident throughput ----
    pr.reset()
    for i in range(args.warmup):
        step()
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    kev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    launches0 = ctx.get_stat("trace_launches")
    for k in range(args.steps):
        flush.fill_(k & 0xFF)                   # L2 flush between timed iterations (not timed)
        barrier()
        kev_cur[0] = kev[k]
        ev[k][0].record()
        step(last=(k == args.steps - 1))        # the K timed steps are one frame of K * S * N samples: one NCCL reduce at its end
        ev[k][1].record()
        kev_cur[0] = None
        barrier()
    launches = ctx.get_stat("trace_launches") - launches0      # counted by the library at its launch sites
    clocks = sampler.stop() if rank == 0 else None
    ms = sum(a.elapsed_time(b) for a, b in ev)
    kms = sum(a.elapsed_time(b) for a, b in kev) / args.steps
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total = float(t.item())
    samples_per_step = W * H * S * world
    value = samples_per_step * args.steps / (ms_total * 1e-3)

    # ---- end to end through the C ABI with host buffers ----
    # pinned host buffers for everything that crosses PCIe inside the timed region
    pinned = []

    def pin(arr):
        t = torch.from_numpy(np.ascontiguousarray(arr)).pin_memory()
        pinned.append(t)
        return t.numpy()

    import copy
    grid = copy.copy(grid)
    grid.indirection, grid.range, grid.atlas, grid.mips = pin(grid.indirection), pin(grid.range), pin(grid.atlas), [pin(m) for m in grid.mips]
    env, lut = pin(env), pin(lut)
    h2d = grid.indirection.nbytes + grid.range.nbytes + grid.atlas.nbytes + sum(m.nbytes for m in grid.mips) + env.nbytes + lut.nbytes
    d2h = H * W * 16 if rank == 0 else 0

    # Four contexts on this GPU, each with its own stream, colour buffer and pinned read-back image (a ring of in-flight
    # frames through the public API): step k runs on lane k % 4 with asynchronous uploads ("async_upload": the pinned inputs
    # outlive the step), so the host enqueues upload + trace of the next steps while it waits for the device->host read of
    # an older one, and the tracking kernels run back to back. Measured on B200: 2 lanes with blocking uploads 21.6 G,
    # 2 / 3 / 4 lanes with asynchronous uploads 20.9 / 25.1 / 29.5 Gsamples/s (the persistent tracking kernel fills every
    # SM, so another lane's small upload kernels only run between two traces; a blocking upload stalls the host on them).
    # Every step still uploads its own inputs and its image is read back inside the timed region.
    class Lane:
        def __init__(self):
            self.stream = torch.cuda.Stream(device=dev)
            self.ctx = vr.Context(local_rank)
            self.ctx.set_stream(self.stream.cuda_stream)
            self.ctx.resize(W, H)
            self.ctx.set_option("async_upload", 1)      # the pinned inputs outlive every step: uploads only enqueue
            self.color = torch.zeros((H, W, 4), dtype=torch.float32, device=dev)
            self.ctx.bind_color(self.color.data_ptr())
            self.host_img = pin(np.empty((H, W, 4), np.float32))
            self.pending = False
            lane_ctx = self.ctx
            self.pr = PartitionedRenderer(self.color, lambda first, n, tile, accum: lane_ctx.trace(params, first, n, tile=tile, accum_mode=accum), partition="spp")

        def finish(self):
            if self.pending:
                if rank == 0:
                    self.ctx.lib.vrb_download_color(self.ctx.handle, self.host_img.ctypes.data, 4)   # device -> host (blocks on this lane's stream)
                else:
                    self.ctx.sync()
                self.pending = False

        def submit(self):
            with torch.cuda.stream(self.stream):
                self.ctx.grid_upload_brick(grid)        # host -> device: indirection, range, atlas, mips
                self.ctx.env_upload(env)                # host -> device + importance pyramid rebuild
                self.ctx.tf_upload(lut)
                self.pr.render(S * world)
            self.pending = True

    lanes = [Lane() for _ in range(max(2, int(os.environ.get("VRB_E2E_LANES", "4"))))]
    n_lanes = len(lanes)
    # the pipeline needs a few steps to fill and one read-back to drain: time at least 40 steps so that the number is the
    # steady state whatever --steps is (reported as e2e.steps)
    e2e_steps = max(args.steps, 40)

    def e2e_step(k):
        lane = lanes[k % n_lanes]
        lane.finish()                                   # the image of step k - 2 (normally long done)
        lane.submit()

    for k in range(4):
        e2e_step(k)
    for lane in lanes:
        lane.finish()
    barrier()
    t0 = time.perf_counter()
    for k in range(e2e_steps):
        e2e_step(k)
    for lane in lanes:
        lane.finish()
    barrier()
    e2e_s = time.perf_counter() - t0
    t = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_value = samples_per_step * e2e_steps / float(t.item())

    if rank == 0:
        peaks, peak_kind = measured_peaks()
        traffic, traffic_src = None, "no capture committed"
        try:        # DRAM bytes of one launch of the same kernel on the same workload, from the committed ncu capture
            tj = json.load(open(os.path.join(ROOT, "profiles", "traffic_latest.json")))
            traffic, traffic_src = tj["dram_read_bytes"] + tj["dram_write_bytes"], tj["source"]
        except Exception:
            pass
        achieved = alg_bytes / (kms * 1e-3) / 1e9
        out = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "reference assets (smoke.brick, lut.txt, hdr) committed under tests/golden/assets",
            "config": {
                "workload": "configs[1]: smoke.brick + lut.txt TF (pathtracer_brick_tf), 1920x1080, 128 bounces",
                "spp_per_step": S, "full_config_spp": FULL_SPP, "partition": "spp slices, one NCCL reduce(SUM) of the float4 buffers at the end of the K-step frame (inside the timed region)" if world > 1 else "single GPU",
                "l2": "flushed between timed iterations (256 MiB fill); the 1.9 MB volume is L2-resident by nature",
            },
            "roofline": {
                "bound": "hbm", "achieved": achieved, "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": achieved / peaks["hbm_gbs"],
                "traffic": traffic, "traffic_unit": "bytes per launch (ncu dram__bytes_read.sum + dram__bytes_write.sum, " + traffic_src + ")",
                "algorithmic_bytes_per_launch": alg_bytes, "peak_kind": peak_kind, "kernel": "k_trace_persistent<TF, FastMath>", "kernel_ms": kms,
                "algorithmic_bytes_per_sample": alg_bytes / counters["n_samples"], "counters_per_sample": {k: v / counters["n_samples"] for k, v in counters.items()},
                "executed": {"bytes_per_sample": exec_bytes / counters["n_samples"], "achieved": exec_bytes / (kms * 1e-3) / 1e9,
                             "traced_fraction": counters_exec["n_samples"] / counters["n_samples"],
                             "note": "events the production launch executes after its exact screen-space culling; `achieved` above uses the reference algorithm's bytes for the same image"},
                "note": "latency/issue-bound gather workload on an L2-resident volume: see DESIGN.md for the L2 roofline",
            },
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h), "steps": e2e_steps, "lanes": n_lanes},
            "gpu_launches": int(launches),      # rank 0's own kernels in the timed region (tracking kernel + k_fold per step, brick mask + tile keys when the cached order is rebuilt)
            "clocks": clocks,
        }
        if world == 1 and not args.no_cpu_baseline:
            # separate process: no torch / CUDA runtime (and no second OpenMP runtime) next to the OpenMP oracle
            r = subprocess.run([sys.executable, os.path.abspath(__file__), "--cpu-baseline-worker"], capture_output=True, text=True)
            try:
                out["cpu_baseline"] = json.loads(r.stdout.strip().splitlines()[-1])
            except Exception:
                out["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": None, "kind": "port", "sample": "failed: " + (r.stderr or "")[-300:]}
        print(json.dumps(out))
    for lane in lanes:
        lane.ctx.close()
    if world > 1:
        dist.destroy_process_group()
    ctx.close()


if __name__ == "__main__":
    main()
