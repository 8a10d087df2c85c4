#!/usr/bin/env python
"""bench.py -- path samples/s of the B200 volume path tracer on BASELINE.json's configs.

    python bench.py --gpus N --steps K --warmup W [--impl reference] [--configs C1,C3,C4,C5 | none]

Headline workload at every N: BASELINE.json configs[1] (C2) -- data/smoke.brick + data/lut.txt transfer function
(pathtracer_brick_tf), 1920x1080, 128 bounces, 1024 spp. One STEP = one full frame of the config on every GPU
(--spp-per-step = 1024 samples per pixel and rank; internally 32 launches of the tracking kernel + fold of 32 spp each).

value     whole-job samples/s with volume/env/LUT resident in HBM: CUDA events around the K steps, max over ranks.
          N > 1: weak scaling -- every rank traces its own 1024-spp slice of sample indices of an N*1024-spp frame into a SUM
          buffer and EVERY step ends with the one exchange of the path, the NCCL reduce(SUM) of the float4 images to rank 0
          (inside the timed region; its own duration is reported as collective_ms).
e2e       the same metric through the C ABI with HOST buffers: per step (= per frame) the brick grid, environment and LUT are
          uploaded from pinned host memory, the frame is traced and the RGBA32F image is read back, all inside the timed region.
roofline  headline kernel (C2 is L2-resident by nature: 27 MB working set): EXECUTED algorithmic bytes (event counters of the
          launch the timed region runs) / kernel time against the L2 random-sector-gather ceiling measured in
          this run (vrb_probe_bandwidth); the HBM numbers are next to it. C3 / C4 (HBM-resident atlases) report against HBM.
configs   C1, C3, C4 (samples/s, roofline, counters) and C5 (frames/s) measured in the same run; at N > 1 they are split the way
          BASELINE.json names (spp slices + reduce per frame for C1/C3/C4, frames dealt to the ranks for C5): strong scaling.
strong    N > 1: ONE 1024-spp frame of the headline config split over the N ranks with its per-frame reduce.
cpu_baseline / --impl reference: the reference's OWN tracking code on all host threads -- its unmodified GLSL compiled as C++
          (oracle/_ref/libglsl_ref.so, kind "reference"; the restated port oracle/vr_oracle.c, bit-identical to it, only where
          that library is absent, kind "port") -- on a bounded sample of the same frame (480x270, same camera).
"""
from __future__ import annotations

import argparse
import glob
import hashlib
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
sys.path.insert(0, os.path.join(ROOT, "tools"))

W, H, BOUNCES, FULL_SPP = 1920, 1080, 128, 1024
METRIC = "path samples/sec (1080p, 128 bounces)"
UNIT = "samples/s"
WORKLOAD = "configs[1]: smoke.brick + lut.txt TF (pathtracer_brick_tf), 1920x1080, 128 bounces, 1024 spp"
CPU_W, CPU_H = W // 4, H // 4     # CPU legs: the SAME frame (camera, fov, aspect) at 1/4 resolution per axis


def load_workload(w=W, h=H):
    import workloads as wl
    grid, env, lut = wl.load_assets()
    # `./volren data/smoke.brick <hdr> data/lut.txt -w 1920 -h 1080 --render --bounces 128`
    return grid, env, lut, wl.default_params(grid, w, h, bounces=BOUNCES, use_tf=True)


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return json.load(open(p)), "measured (MEASURED_PEAKS.json)"
    return {"hbm_gbs": 6650.0}, "fallback (B200_PROFILING.md)"


def source_hash():
    """sha256 over the CUDA sources: ncu captures under profiles/ are stamped with it and ignored when the kernels changed."""
    h = hashlib.sha256()
    for f in sorted(glob.glob(os.path.join(ROOT, "volren_b200", "csrc", "*.cu*"))):
        h.update(os.path.basename(f).encode())
        h.update(open(f, "rb").read())
    return h.hexdigest()[:16]


def committed_traffic(key):
    """DRAM bytes per launch of the tracking kernel on workload `key` from the committed ncu capture -- only when the capture
    was taken from the sources this run uses."""
    try:
        tj = json.load(open(os.path.join(ROOT, "profiles", "traffic_latest.json")))
    except Exception:
        return None, "no capture committed"
    if tj.get("source_sha16") != source_hash():
        return None, f"stale capture ignored (taken at source hash {tj.get('source_sha16')}, this run is {source_hash()})"
    e = tj.get(key)
    if not e:
        return None, "no capture for this workload"
    return e["dram_read_bytes"] + e["dram_write_bytes"], e.get("source", "profiles/traffic_latest.json")


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
            import torch
            try:        # CUDA_VISIBLE_DEVICES may remap indices: match by PCI bus id of the torch device when possible
                pr = torch.cuda.get_device_properties(index)
                self.handle = pynvml.nvmlDeviceGetHandleByPciBusId(f"{getattr(pr, 'pci_domain_id', 0):08x}:{pr.pci_bus_id:02x}:{getattr(pr, 'pci_device_id', 0):02x}.0".encode())
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


# ---------------------------------------------------------------------------------------------------------------------
# CPU legs

def _cpu_trace_rate(spp_budget_s, steps, warmup):
    """The reference's shaders on the host cores, on the headline frame at 480x270 -> (samples/s, cores, kind, sample, dt per step).
    kind "reference": oracle/_ref/libglsl_ref.so -- the reference's UNMODIFIED pathtracer_brick_tf.glsl + common.glsl compiled
    as C++ over the reference's own glm (oracle/Makefile; OpenMP over the rows of the dispatch grid), prebuilt where
    /root/reference exists and shipped to the GPU box; kind "port": the restatement oracle/vr_oracle.c when that library is
    absent (the two are bit-identical, tests/test_glsl_ref.py; the compiled shaders are ~3x faster than the port)."""
    from oracle.binding import GlslRef, Oracle
    grid, env, lut, params = load_workload(CPU_W, CPU_H)
    o = Oracle()
    pyr = o.env_build(env)
    sc = o.make_scene(grid, env, pyr, lut=lut)
    cores = len(os.sched_getaffinity(0))     # all host threads, whatever OMP_NUM_THREADS torchrun exported
    color = np.zeros((CPU_H, CPU_W, 4), np.float32)
    if GlslRef.available():
        g = GlslRef()
        kind, what = "reference", "oracle/_ref/libglsl_ref.so = the reference's own GLSL compiled as C++, OpenMP"
        run = lambda first, n: g.trace(sc, params, first, n, color=color, n_threads=cores)
    else:
        kind, what = "port", "oracle/vr_oracle.c, OpenMP"
        run = lambda first, n: o.trace(sc, params, first, n, color=color, n_threads=cores)
    run(1, 1)
    t0 = time.perf_counter()
    run(1, 4)
    probe = (time.perf_counter() - t0) / 4
    spp = max(1, int(spp_budget_s / max(probe, 1e-4)))
    for i in range(warmup):
        run(1 + i, 1)
    t0 = time.perf_counter()
    for k in range(steps):
        run(1 + k * spp, spp)
    dt = time.perf_counter() - t0
    sample = f"the 1080p frame rendered at {CPU_W}x{CPU_H} (same camera), {spp} spp per step, {steps} step(s); {what}"
    if kind == "reference":      # for transparency: the hand-restated port (vr_oracle.c) on the same frame, one short run
        n_port = 256
        t1 = time.perf_counter()
        o.trace(sc, params, 1, n_port, color=color, n_threads=cores)
        sample += f"; the restated port oracle/vr_oracle.c runs the same frame at {CPU_W * CPU_H * n_port / (time.perf_counter() - t1):.3g} samples/s ({n_port} spp)"
    return CPU_W * CPU_H * spp * steps / dt, cores, kind, sample, dt / steps


def run_reference(args, rank, world):
    """--impl reference: the reference's CPU implementation of the path. The reference's renderer needs an OpenGL 4.5 context
    that neither this container nor the GPU box has (profiles/r02_gl_probe.txt), so its shaders run as compiled C++ on all host
    threads (kind = "reference", see _cpu_trace_rate); each step is a bounded sample of the same frame."""
    if rank != 0:
        return
    v, cores, kind, sample, dt = _cpu_trace_rate(min(3.0, 90.0 / max(args.steps, 1)), args.steps, args.warmup)
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * dt, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "reference assets (smoke.brick, lut.txt, hdr) committed under tests/golden/assets",
        "config": {"workload": WORKLOAD, "sample": sample},
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def cpu_baseline_worker():
    v, cores, kind, sample, _ = _cpu_trace_rate(12.0, 1, 0)
    print(json.dumps({"value": v, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample}))


# ---------------------------------------------------------------------------------------------------------------------
# GPU arm

def main():
    if "--cpu-baseline-worker" in sys.argv:
        cpu_baseline_worker()
        return
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--spp-per-step", type=int, default=FULL_SPP, help="samples per pixel, rank and step (default: the config's full 1024-spp frame)")
    ap.add_argument("--configs", default="C1,C3,C4,C5", help="other BASELINE configs measured into the `configs` object ('none' to skip)")
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
    import workloads as wl
    from volren_b200.multigpu import PartitionedRenderer

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
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)   # > 126 MB L2

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def ev_pair():
        return torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)

    def max_over_ranks(x):
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def counters_of(c, p, spp, culled):
        """event counters of one launch of `spp` samples per pixel (the counting build of the same kernel)"""
        c.set_option("count_culled", 1 if culled else 0)
        c.set_counting(True)
        c.clear()
        c.trace(p, 1, spp)
        out = c.get_counters().as_dict()
        c.set_counting(False)
        c.set_option("count_culled", 0)
        c.clear()
        return out

    class Timed:
        """Runs `frames` of `spp_total` samples on a context partitioned over the ranks by spp slices, one reduce per frame."""

        def __init__(self, c, p, w, h):
            self.c, self.p, self.w, self.h = c, p, w, h
            c.resize(w, h)
            self.color = torch.zeros((h, w, 4), dtype=torch.float32, device=dev)
            c.bind_color(self.color.data_ptr())
            self.kev = None
            self.coll = []
            self.pr = PartitionedRenderer(self.color, self._trace, partition="spp")

        def _trace(self, first, n, tile, accum):
            if self.kev is not None:
                self.kev[0].record()
            self.c.trace(self.p, first, n, tile=tile, accum_mode=accum)
            if self.kev is not None:
                self.kev[1].record()

        def frame(self, spp_total, kev=None):
            """one frame: spp_total samples split over the ranks, then the reduce (timed separately as the collective)"""
            self.pr.reset()
            self.kev = kev
            self.pr.render(spp_total, reduce=False)
            self.kev = None
            if world > 1:
                a, b = ev_pair()
                a.record()
                self.pr.finish()
                b.record()
                self.coll.append((a, b))

        def run(self, spp_total, frames, warm_spp):
            """-> (samples/s whole job, ms per frame (max over ranks), kernel ms per frame on this rank, collective ms per frame)"""
            self.frame(warm_spp)
            barrier()
            self.coll = []
            launches0 = self.c.get_stat("trace_launches")
            ev, kev = [ev_pair() for _ in range(frames)], [ev_pair() for _ in range(frames)]
            for k in range(frames):
                flush.fill_(k & 0xFF)               # L2 flush between timed iterations (not timed)
                barrier()
                ev[k][0].record()
                self.frame(spp_total, kev[k])
                ev[k][1].record()
                barrier()
            self.launches = self.c.get_stat("trace_launches") - launches0     # this rank's own kernels inside the timed region
            ms = max_over_ranks(sum(a.elapsed_time(b) for a, b in ev))
            kms = sum(a.elapsed_time(b) for a, b in kev) / frames
            cms = sum(a.elapsed_time(b) for a, b in self.coll) / max(len(self.coll), 1) if self.coll else 0.0
            return self.w * self.h * spp_total * frames / (ms * 1e-3), ms / frames, kms, cms

    # ---- ceilings measured in this run, on this device ("must be micro-benchmarked on the box") ----
    peaks, peak_kind = measured_peaks()
    probes = None
    if rank == 0:
        probes = {"l2_gather_gbs": ctx.probe_bandwidth(32 << 20, 1), "l2_stream_gbs": ctx.probe_bandwidth(32 << 20, 0),
                  "hbm_gather_gbs": ctx.probe_bandwidth(1 << 30, 1), "hbm_stream_gbs": ctx.probe_bandwidth(1 << 30, 0),
                  "how": "vrb_probe_bandwidth: 32 MiB (L2-resident) and 1 GiB working sets; random 32-B sector gathers counted at 32 B each, 16-B streaming loads; best of 5"}

    # ---- headline: C2 resident in HBM ----
    ctx.grid_upload_brick(grid)
    ctx.env_upload(env)
    ctx.tf_upload(lut)
    head = Timed(ctx, params, W, H)
    PASS = 32
    counters = counters_of(ctx, params, PASS, culled=False)         # the reference algorithm's events for this image
    counters_exec = counters_of(ctx, params, PASS, culled=True)     # the events the production launch executes (exact culling kept)
    n_all = counters["n_samples"]
    alg_bytes = wl.algorithmic_bytes(counters, True)
    exec_bytes = wl.algorithmic_bytes(counters_exec, True) + 32 * (n_all - counters_exec["n_samples"])   # folded zeros still cost the RMW

    for _ in range(max(0, args.warmup - 1)):        # W warm-up steps in total: run() does the last one itself
        head.frame(S * world)
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    value, ms_per_step, kms, coll_ms = head.run(S * world, args.steps, S * world)
    launches = head.launches
    clocks = sampler.stop() if rank == 0 else None
    launches_per_step = -(-S // PASS)
    k_launch_ms = kms / launches_per_step                            # tracking kernel + its fold, per 32-spp launch

    # ---- strong scaling of the headline frame: 1024 spp split over the ranks, reduce per frame ----
    strong = None
    if world > 1:
        sv, sms, skms, scms = head.run(FULL_SPP, max(3, min(args.steps, 10)), FULL_SPP)
        strong = {"value": sv, "unit": UNIT, "ms_per_frame": sms, "kernel_ms_per_frame": skms, "collective_ms": scms,
                  "frame": f"{FULL_SPP} spp split into {world} slices of {FULL_SPP // world}, one NCCL reduce(SUM) of the 33 MB float4 image per frame"}

    # ---- end to end through the C ABI with host buffers ----
    pinned = []

    def pin(arr):
        t = torch.from_numpy(np.ascontiguousarray(arr)).pin_memory()
        pinned.append(t)
        return t.numpy()

    import copy
    hgrid = copy.copy(grid)
    hgrid.indirection, hgrid.range, hgrid.atlas, hgrid.mips = pin(grid.indirection), pin(grid.range), pin(grid.atlas), [pin(m) for m in grid.mips]
    henv, hlut = pin(env), pin(lut)
    h2d = hgrid.indirection.nbytes + hgrid.range.nbytes + hgrid.atlas.nbytes + sum(m.nbytes for m in hgrid.mips) + henv.nbytes + hlut.nbytes
    d2h = H * W * 16 if rank == 0 else 0

    # Two contexts on this GPU, each with its own stream, colour buffer and pinned read-back image: step k runs on lane k % 2
    # with asynchronous uploads (the pinned inputs outlive the step), so the host enqueues the next frame while it waits for
    # the read-back of the previous one. Every step uploads its own inputs and its image is read back inside the timed region.
    class Lane:
        def __init__(self):
            self.stream = torch.cuda.Stream(device=dev)
            self.ctx = vr.Context(local_rank)
            self.ctx.set_stream(self.stream.cuda_stream)
            self.ctx.resize(W, H)
            self.ctx.set_option("async_upload", 1)
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
                self.ctx.grid_upload_brick(hgrid)       # host -> device: indirection, range, atlas, mips
                self.ctx.env_upload(henv)               # host -> device + importance pyramid rebuild
                self.ctx.tf_upload(hlut)
                self.pr.reset()
                self.pr.render(S * world)               # incl. the per-frame reduce at N > 1
            self.pending = True

    lanes = [Lane() for _ in range(max(2, int(os.environ.get("VRB_E2E_LANES", "2"))))]
    e2e_steps = max(args.steps, 8)
    for k in range(len(lanes)):
        lanes[k].submit()
    for lane in lanes:
        lane.finish()
    barrier()
    t0 = time.perf_counter()
    for k in range(e2e_steps):
        lane = lanes[k % len(lanes)]
        lane.finish()
        lane.submit()
    for lane in lanes:
        lane.finish()
    barrier()
    e2e_s = max_over_ranks(time.perf_counter() - t0)
    e2e_value = W * H * S * world * e2e_steps / e2e_s
    for lane in lanes:
        lane.ctx.close()
    del lanes

    # ---- the other BASELINE configs ----
    cfg_out = {}
    wanted = [] if args.configs.lower() == "none" else [c.strip().upper() for c in args.configs.split(",") if c.strip()]

    def roofline_of(c, p, tf, kernel_ms_per_launch, spp_launch, bound, traffic_key):
        cs = counters_of(c, p, 2, culled=False)
        ce = counters_of(c, p, 2, culled=True)
        n = cs["n_samples"]
        per_sample = wl.algorithmic_bytes(ce, tf) / n + 32.0 * (n - ce["n_samples"]) / n
        bytes_launch = per_sample * p.resolution[0] * p.resolution[1] * spp_launch
        achieved = bytes_launch / (kernel_ms_per_launch * 1e-3) / 1e9
        peak = probes["l2_gather_gbs"] if bound == "l2" else peaks["hbm_gbs"]
        traffic, src = committed_traffic(traffic_key)
        return {"bound": bound, "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "peak_kind": "L2 random 32-B sector gather ceiling measured in this run (vrb_probe_bandwidth)" if bound == "l2" else peak_kind,
                "traffic": traffic, "traffic_source": src, "algorithmic_bytes_per_sample": per_sample,
                "reference_algorithm_bytes_per_sample": wl.algorithmic_bytes(cs, tf) / n, "traced_fraction": ce["n_samples"] / n,
                "counters_per_sample": {k: v / n for k, v in cs.items()}}

    def run_config(name):
        c2 = vr.Context(local_rank)
        c2.set_stream(stream.cuda_stream)
        c2.env_upload(env)
        try:
            if name == "C1":
                w, h, spp, tf, bound = 1024, 1024, 64, False, "l2"
                c2.grid_upload_brick(grid)
                p = wl.readme_params(grid, w, h)
                label, extra = "configs[0]: smoke.brick + hdr, README command (non-TF, environment visible), 1024x1024, 64 spp", {}
            else:
                w, h = W, H
                if name == "C3":
                    n = 1024
                    vox, dims, tf, spp, bound = wl.fbm_cloud(n), (n, n, n), False, 4096, "hbm"
                    label = "configs[2]: synthetic 1024^3 fBm cloud -> GPU brick build, density 100, albedo .8 (non-TF), 1920x1080, 4096 spp"
                else:
                    vox, dims, tf, spp, bound = wl.ct_phantom(512, 512, 1800), (512, 512, 1800), True, 1024, "hbm"
                    c2.tf_upload(wl.turbo_lut())
                    label = "configs[3]: synthetic 512x512x1800 CT phantom + 256-entry Turbo LUT (TF), 1920x1080, 1024 spp"
                torch.cuda.synchronize()
                torch.cuda.empty_cache()
                c2.grid_build_from_dense_device(vox.data_ptr(), dims, 0.0, 1.0)       # first build grows the memory pool
                a, b = ev_pair()
                a.record()
                c2.grid_build_from_dense_device(vox.data_ptr(), dims, 0.0, 1.0)
                b.record()
                torch.cuda.synchronize()
                build_ms = a.elapsed_time(b)
                nb, _, count = c2.grid_info()
                n_vox = dims[0] * dims[1] * dims[2]
                build_bytes = n_vox + count * 512 + 8 * nb[0] * nb[1] * nb[2]
                extra = {"grid": list(dims), "n_bricks": list(nb), "bricks_allocated": int(count), "atlas_MiB": count * 512 / 2 ** 20,
                         "brick_build": {"ms": build_ms, "algorithmic_GBps": build_bytes / (build_ms * 1e-3) / 1e9, "peak": peaks["hbm_gbs"],
                                         "frac": build_bytes / (build_ms * 1e-3) / 1e9 / peaks["hbm_gbs"],
                                         "bytes": "1 B/voxel read + 1 B/allocated voxel written + 8 B/brick (SURVEY 8(d))"}}
                if name == "C3":
                    # the pipeline SURVEY 8(d) names for C3: fp32 field -> DenseGrid(float*) (global min/max + 8-bit quantisation)
                    # -> BrickGrid, all on the GPU (vrb_grid_build_from_float_device). The u8 codes of `vox` as floats give the
                    # same DenseGrid (min 0, max 255 -> the identical codes), so the grid the frame is traced on does not change.
                    f32 = vox.float()
                    torch.cuda.synchronize()
                    c2.grid_build_from_float_device(f32.data_ptr(), dims, frame=1)
                    a, b = ev_pair()
                    a.record()
                    mm = c2.grid_build_from_float_device(f32.data_ptr(), dims, frame=1)
                    b.record()
                    torch.cuda.synchronize()
                    f_ms = a.elapsed_time(b)
                    f_bytes = 9 * n_vox + build_bytes
                    extra["dense_from_float_and_brick_build"] = {
                        "ms": f_ms, "algorithmic_GBps": f_bytes / (f_ms * 1e-3) / 1e9, "frac": f_bytes / (f_ms * 1e-3) / 1e9 / peaks["hbm_gbs"],
                        "min_max": list(mm), "bytes": "DenseGrid(float*): 4 B/voxel read twice + 1 B/voxel written, then the brick build's bytes"}
                    c2.lib.vrb_grid_free(c2.handle, 0, 1)
                    del f32
                del vox
                torch.cuda.empty_cache()
                p = wl.synthetic_params(dims, w, h, tf)
            t = Timed(c2, p, w, h)
            v, ms, kms_c, cms = t.run(spp, 1 if spp >= 1024 else 4, min(spp, 64))
            out = {"workload": label, "value": v, "unit": UNIT, "resolution": [w, h], "spp": spp, "ms_per_frame": ms, "collective_ms": cms,
                   "partition": f"{world} spp slices + one NCCL reduce per frame" if world > 1 else "single GPU", **extra}
            if rank == 0:
                per_rank_spp = spp // world if world > 1 else spp
                out["roofline"] = roofline_of(c2, p, tf, kms_c / max(1, -(-per_rank_spp // PASS)), min(PASS, per_rank_spp), bound, name.lower())
            return out
        finally:
            c2.close()

    def run_c5(frames=8, clean_spp=256, n=256, w=1024, h=1024):
        """datagen_denoise.py:60-130 on synthetic animated frames: per frame GPU brick build, noisy + clean render, two read-backs
        to fp16 (N, 3, H, W); frames are dealt round-robin to the ranks (independent jobs: no exchange)."""
        c5 = vr.Context(local_rank)
        c5.set_stream(stream.cuda_stream)
        c5.env_upload(env)
        c5.resize(w, h)
        vols = wl.fbm_frames(n, frames)
        plist = wl.c5_parameters(frames)
        inputs = np.zeros((frames, 3, h, w), np.float16)
        targets = np.zeros((frames, 3, h, w), np.float16)
        samples = 0
        barrier()
        t0 = time.perf_counter()
        for i, q in enumerate(plist):
            if i % world != rank:
                continue
            c5.grid_build_from_dense_device(vols[i].data_ptr(), (n, n, n), 0.0, 1.0)
            for seed, spp, dst in ((q["seed_input"], q["samples"], inputs), (q["seed_target"], clean_spp, targets)):
                c5.clear()
                c5.trace(wl.c5_frame_params(q, n, w, h, seed), 1, spp)
                rgb = c5.download_color(3)                                             # fbo_data(): linear RGB fp32
                dst[i] = np.transpose(np.flip(rgb, axis=0).astype(np.float16), [2, 1, 0])  # datagen_denoise.py:113-114
                samples += w * h * spp
        torch.cuda.synchronize()
        dt = max_over_ranks(time.perf_counter() - t0)
        tot = torch.tensor([float(samples)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(tot, op=dist.ReduceOp.SUM)
        c5.close()
        return {"workload": f"configs[4]: {frames} animated {n}^3 fBm frames, datagen_denoise-style pairs (noisy 1..33 spp, clean {clean_spp} spp; named: 64 frames x 4096 spp), 1024x1024, fp16 (N,3,H,W) out",
                "value": frames / dt, "unit": "frames/s", "samples_per_s": float(tot.item()) / dt, "wall_s": dt,
                "partition": f"frames dealt round-robin to {world} ranks, no exchange" if world > 1 else "single GPU",
                "finite": bool(np.isfinite(inputs.astype(np.float32)).all() and np.isfinite(targets.astype(np.float32)).all())}

    for name in wanted:
        try:
            cfg_out[name] = run_c5() if name == "C5" else run_config(name)
        except Exception as e:      # a failing side config must not take the headline line down
            cfg_out[name] = {"error": f"{type(e).__name__}: {e}"[:300]}

    if rank == 0:
        traffic, traffic_src = committed_traffic("c2")
        achieved = exec_bytes / (k_launch_ms * 1e-3) / 1e9       # executed algorithmic bytes of one 32-spp launch / its duration
        l2_peak = probes["l2_gather_gbs"]
        out = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "reference assets (smoke.brick, lut.txt, hdr) committed under tests/golden/assets; synthetic grids (seed 42) for C3-C5",
            "config": {
                "workload": WORKLOAD, "spp_per_step": S, "full_config_spp": FULL_SPP,
                "step": "one full frame of the config per GPU (32 launches of the tracking kernel + fold, 32 spp each)",
                "partition": (f"weak scaling: {world} spp slices of {S} spp, one NCCL reduce(SUM) of the float4 images per step (inside the timed region)" if world > 1 else "single GPU"),
                "l2": "flushed between timed iterations (256 MiB fill); the 1.9 MB volume / 27 MB decoded working set is L2-resident by nature",
            },
            "roofline": {
                "bound": "l2", "achieved": achieved, "peak": l2_peak, "unit": "GB/s", "frac": achieved / l2_peak,
                "peak_kind": "L2 random 32-B sector gather ceiling measured in this run (vrb_probe_bandwidth, 32 MiB working set)",
                "hbm": {"peak": peaks["hbm_gbs"], "peak_kind": peak_kind, "frac": achieved / peaks["hbm_gbs"],
                        "note": "C2's working set never leaves L2: DRAM moves <1 % of the algorithmic bytes, so HBM does not bound this workload; C3 / C4 in `configs` are the HBM-resident ones"},
                "traffic": traffic, "traffic_unit": "bytes per launch (ncu dram__bytes_read.sum + dram__bytes_write.sum)", "traffic_source": traffic_src,
                "algorithmic_bytes_per_launch": exec_bytes, "algorithmic_bytes_per_sample": exec_bytes / n_all,
                "bytes_definition": "EXECUTED events of one 32-spp launch (counting build with the production culling kept) x SURVEY 8(d) bytes per event; culled samples still pay their 32 B accumulate",
                "reference_algorithm_bytes_per_sample": alg_bytes / n_all, "traced_fraction": counters_exec["n_samples"] / n_all,
                "kernel": "k_trace_pool<TF, FastMath> + k_fold", "kernel_ms": k_launch_ms, "launches_per_step": launches_per_step,
                "counters_per_sample": {k: v / n_all for k, v in counters.items()},
                "probes": probes,
            },
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h), "steps": e2e_steps,
                    "note": "per step: upload grid + env + LUT from pinned host memory, trace the frame, read the RGBA32F image back"},
            "collective_ms": coll_ms,
            "gpu_launches": int(launches),      # rank 0's own kernels in the timed region (tracking kernel + k_fold per pass; brick mask + tile keys when the cached order is rebuilt)
            "clocks": clocks,
            "configs": cfg_out,
            "source_sha16": source_hash(),
        }
        if strong is not None:
            out["strong"] = strong
        if world == 1 and not args.no_cpu_baseline:
            # separate process: no torch / CUDA runtime (and no second OpenMP runtime) next to the OpenMP oracle
            r = subprocess.run([sys.executable, os.path.abspath(__file__), "--cpu-baseline-worker"], capture_output=True, text=True)
            try:
                out["cpu_baseline"] = json.loads(r.stdout.strip().splitlines()[-1])
            except Exception:
                out["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": None, "kind": "port",
                                       "sample": f"failed: rc {r.returncode} " + (r.stderr or r.stdout or "")[-300:]}
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()
    ctx.close()


if __name__ == "__main__":
    main()
