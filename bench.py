#!/usr/bin/env python
"""bench.py -- Mrays/s (primary + secondary) of the ray-cast + radiance path on B200.

  python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path (through the C-ABI)
  python bench.py --impl reference --gpus N --steps K ...  # the reference algorithm on the host cores

Workload (BASELINE.json configs[1], the config the metric is quoted on): cornellbox.bin ambient
occlusion, 1024x1024, 16 AO rays per pixel.  One STEP = one pass of the hot path over one frame of
that image (1 primary + 16 AO rays per pixel = 17.8 Mrays) per rank.  With N ranks the frames are
dealt round-robin (rank r renders frame step*N + r: per-GPU work is fixed -> "weak" scaling, no
data-path collective); the per-rank linear accumulators are combined ONCE with an NCCL reduce at the
end of the timed region.  value = rays of all ranks / max-over-ranks device time.

Printed keys follow the driver's contract; see DESIGN.md "Measurement" for how each number is made.
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
SCENE = os.path.join(ROOT, "data", "cornellbox.bin")

WORKLOADS = {
    # name: (width, height, mode, params)
    "c1": dict(width=512, height=512, mode=0, desc="cornellbox.bin primary-ray cast 512x512, 1 spp, hit-ID output"),
    "c2": dict(width=1024, height=1024, mode=1, ao_samples=16, desc="cornellbox.bin ambient occlusion 1024x1024, 16 AO rays/pixel"),
    "c3": dict(width=1920, height=1080, mode=2, desc="cornellbox.bin direct lighting with area-light shadow rays at 1920x1080"),
    "c4": dict(width=3840, height=2160, mode=3, max_depth=8, desc="cornellbox.bin full path trace (max depth 8) at 4K"),
    "c5": dict(width=3840, height=2160, mode=3, max_depth=8, tess=236, desc="2M-triangle tessellated cornell box, full path trace (max depth 8) at 4K"),
}


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c2", choices=sorted(WORKLOADS))
    ap.add_argument("--integrator", default="auto", choices=["auto", "mega", "wavefront"])
    ap.add_argument("--accel", default="bvh", choices=["bvh", "brute"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--tune", action="append", default=[], help="experiment knob index=value (ptb_device_set_tuning)")
    ap.add_argument("--fpb", type=int, default=0, help="frames_per_batch override")
    ap.add_argument("--smem-nodes", type=int, default=0, help="BVH nodes staged in shared memory (0 = default)")
    ap.add_argument("--max-leaf", type=int, default=0, help="BVH max triangles per leaf (0 = default)")
    ap.add_argument("--sah-traverse", type=float, default=0.0, help="SAH node-visit cost (0 = default 1.2)")
    ap.add_argument("--force-width", type=int, default=0, help="force the scene form: 1 FLAT, 4 4-wide, 2 binary (0 = the library's pick)")
    ap.add_argument("--sharding", choices=["frames", "image"], default="frames",
                    help="frames (default): rank r renders whole frames = r mod N, weak scaling, one NCCL reduce at the end; "
                         "image: every step is ONE image tile-sharded over the ranks (64-pixel blocks round-robin) with the gather "
                         "fused into the resolve kernel over peer memory, strong scaling")
    ap.add_argument("--spp-per-step", type=int, default=1, help="frames rendered per step (image sharding)")
    ap.add_argument("--gpu-build", action="store_true", help="build the BVH on the device (LBVH) instead of the host SAH builder")
    ap.add_argument("--ab", action="store_true", help="also time the other integrator and brute force (extra keys)")
    return ap.parse_args()


def dist_env():
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    return rank, world, local


def bind_to_gpu_numa(gpu_index):
    """Pins this rank (and so its page-locked host buffers) to the CPUs next to its GPU.  torchrun does not bind
    ranks; with 8 ranks reading 16 MiB frames back every 0.6 ms, buffers on the far socket cap the aggregate
    end-to-end rate.  Returns a short description for config.cpu_affinity."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(gpu_index)
        pynvml.nvmlDeviceSetCpuAffinity(h)
        return "numa-local (nvmlDeviceSetCpuAffinity), %d cpus" % len(os.sched_getaffinity(0))
    except Exception as e:  # plumbing only: never fatal
        return "unbound (%s)" % type(e).__name__


class ClockSampler:
    """nvidia-smi clocks + throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap,timestamp")

    def __init__(self, gpu_index):
        self.rows, self.proc, self.gpu = [], None, gpu_index
        self.t0 = self.t1 = None

    def mark_begin(self):
        self.t0 = time.time()

    def mark_end(self):
        self.t1 = time.time()

    def start(self):
        try:
            import shutil
            cmd = ["nvidia-smi", f"--id={self.gpu}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "20"]
            if shutil.which("stdbuf"):  # line-buffer the pipe: short runs end before a 4 KB block fills
                cmd = ["stdbuf", "-oL"] + cmd
            self.proc = subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except OSError:
            self.proc = None

    def _read(self):
        import datetime
        for line in self.proc.stdout:  # the pipe is block-buffered: use nvidia-smi's own timestamp, not the arrival time
            cols = [c.strip() for c in line.split(",")]
            ts = time.time()
            try:
                ts = datetime.datetime.strptime(cols[9], "%Y/%m/%d %H:%M:%S.%f").timestamp()
            except (IndexError, ValueError):
                pass
            self.rows.append((ts, cols))

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        for _ in range(20):  # a run shorter than nvidia-smi's start-up: take its first rows (just after the timed region)
            if self.rows:
                break
            time.sleep(0.05)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        lo = (self.t0 or 0.0) - 0.03
        hi = (self.t1 or 1e18) + 0.03
        rows = [r for (ts, r) in self.rows if lo <= ts <= hi] or [r for (_, r) in self.rows[-3:]]
        sm = [float(r[1]) for r in rows if len(r) >= 9 and r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in rows if len(r) >= 9 and r[2].replace(".", "").isdigit()]
        pw = [float(r[3]) for r in rows if len(r) >= 9 and r[3].replace(".", "").isdigit()]
        reasons = set()
        for r in rows:
            if len(r) >= 9:
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "samples": len(sm), "reasons": sorted(reasons)}


def load_scene(pt, wl):
    tris, mats = pt.load_model(SCENE)
    light = pt.light_from_quad(tris, 5)
    if wl.get("tess"):
        tris = pt.tessellate(tris, wl["tess"])
    return tris, mats, light


def make_params(pt, wl, **kw):
    p = pt.default_params(width=wl["width"], height=wl["height"], mode=wl["mode"], accum=pt.ACCUM_LINEAR, n_frames=1)
    if "ao_samples" in wl:
        p.ao_samples = wl["ao_samples"]
    if "max_depth" in wl:
        p.max_depth = wl["max_depth"]
    for k, v in kw.items():
        if k in ("light_p1", "light_ea", "light_eb"):
            getattr(p, k)[:] = v
        else:
            setattr(p, k, v)
    return p


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        d = json.load(open(path))
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)", float(d.get("sm_max_mhz", 1965.0))
    return 6650.0, "fallback (B200_PROFILING.md)", 1965.0


# ------------------------------------------------------------------------------------------------------
# reference arm: the reference's algorithm (brute-force GenerateColors restatement) on the host cores
# ------------------------------------------------------------------------------------------------------

def cpu_reference(wl, seconds_budget=12.0, sample_px=None, max_frames=8, n_threads=0):
    """Times the oracle's reference-faithful brute-force path (oracle/: the only CPU code bench.py runs).

    The reference has no CPU executor of its own (ADL's DeviceHost cannot launch kernels, SURVEY.md
    section 0): oracle/_ref/adlTest64 is the genuine test program but it needs an OpenCL *GPU* platform
    (measured once on the B200: profiles/reference_opencl_r01/), so the CPU arm is the port, kind = "port".
    """
    from oracle import binding as ob

    tris, mats = ob.load_model(SCENE)
    p1, ea, eb = ob.light_from_quad(tris, 5)
    if wl.get("tess"):
        tris = ob.tessellate(tris, wl["tess"])
    w, h = wl["width"], wl["height"]
    if sample_px is not None:  # bounded sample: a centred crop is not expressible, so shrink the image
        s = (sample_px / float(w * h)) ** 0.5
        w, h = max(16, int(w * s) // 16 * 16), max(16, int(h * s) // 16 * 16)
    kw = dict(mode=wl["mode"], accum=ob.ACCUM_LINEAR, use_bvh=0, light_p1=p1, light_ea=ea, light_eb=eb, n_threads=n_threads)
    if "ao_samples" in wl:
        kw["ao_samples"] = wl["ao_samples"]
    if "max_depth" in wl:
        kw["max_depth"] = wl["max_depth"]
    rays, t_total, frames = 0, 0.0, 0
    while frames < max_frames and (frames == 0 or t_total < seconds_budget):
        prm = ob.default_params(w, h, first_frame=frames, n_frames=1, **kw)
        t0 = time.perf_counter()
        _, _, ctr = ob.render(prm, tris, mats)
        t_total += time.perf_counter() - t0
        rays += ctr["rays_closest"] + ctr["rays_any"]
        frames += 1
    return {"value": rays / t_total / 1e6, "unit": "Mrays/s", "cores": n_threads or ob.max_threads(), "kind": "port",
            "sample": f"{frames} frame(s) of {w}x{h} ({rays} rays, {t_total:.2f} s), brute force over {len(tris)} triangles as the reference kernel",
            "seconds": t_total, "rays": rays, "frames": frames, "wh": [w, h]}


def run_reference_arm(args, rank, world):
    if rank != 0:
        return
    wl = WORKLOADS[args.workload]
    from oracle import binding as ob

    # one step = one bounded sample: a 512x512 frame of the workload (a quarter of C2's pixels)
    sample_px = min(wl["width"] * wl["height"], 512 * 512) if not wl.get("tess") else 16 * 16
    for _ in range(args.warmup):
        cpu_reference(wl, seconds_budget=0.0, sample_px=sample_px, max_frames=1)
    rays, secs = 0, 0.0
    for _ in range(args.steps):
        r = cpu_reference(wl, seconds_budget=0.0, sample_px=sample_px, max_frames=1)
        rays += r["rays"]
        secs += r["seconds"]
    value = rays / secs / 1e6
    line = {
        "impl": "reference", "metric": "Mrays/sec (primary+secondary)", "value": value, "unit": "Mrays/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": secs / args.steps * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"{args.workload}: {wl['desc']}", "step": f"one {r['wh'][0]}x{r['wh'][1]} frame sample of the workload on the host cores",
                   "algorithm": "reference brute-force loop (GenerateColors.cl:137-154) restated in C, OpenMP over pixels"},
        "cpu_baseline": {"value": value, "unit": "Mrays/s", "cores": ob.max_threads(), "kind": "port", "sample": r["sample"]},
        "e2e": {"value": value, "unit": "Mrays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------------

def main():
    args = parse()
    rank, world, local = dist_env()
    if args.impl == "reference":
        run_reference_arm(args, rank, world)
        return

    import torch
    import torch.distributed as dist

    import oclpathtracer_b200 as pt

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product has no CPU path (use --impl reference for the host baseline)")
    torch.cuda.set_device(local)
    affinity = bind_to_gpu_numa(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    wl = WORKLOADS[args.workload]
    integ = {"auto": pt.INTEGRATOR_AUTO, "mega": pt.INTEGRATOR_MEGAKERNEL, "wavefront": pt.INTEGRATOR_WAVEFRONT}[args.integrator]
    accel = pt.ACCEL_BVH if args.accel == "bvh" else pt.ACCEL_BRUTE

    stream = torch.cuda.current_stream()
    dev = pt.Device(local, stream=stream.cuda_stream)  # enqueue on torch's stream so torch events see the work
    for kv in args.tune:
        i, v = kv.split("=")
        dev.set_tuning(int(i), int(v))
    tris, mats, light = load_scene(pt, wl)
    t0 = time.perf_counter()
    bp = None
    if args.smem_nodes or args.max_leaf or args.sah_traverse or args.force_width:
        bp = pt.bvh_params()
        bp.force_width = args.force_width
        if args.sah_traverse:
            bp.traverse_cost = args.sah_traverse
        if args.smem_nodes:
            bp.smem_nodes = args.smem_nodes
        if args.max_leaf:
            bp.max_leaf = args.max_leaf
    scene = dev.scene(tris, mats, bp, gpu_build=args.gpu_build)
    dev.sync()
    build_s = time.perf_counter() - t0
    npix = wl["width"] * wl["height"]
    frame_t = torch.zeros((npix, 4), dtype=torch.float32, device="cuda")   # this step's mean frame
    accum_t = torch.zeros((npix, 4), dtype=torch.float32, device="cuda")   # linear accumulator over steps
    frame = dev.wrap(frame_t.data_ptr(), frame_t.numel() * 4)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")       # > 126 MB L2

    def params(step_frame, **kw):
        return make_params(pt, wl, first_frame=step_frame, integrator=integ, accel=accel, frames_per_batch=args.fpb,
                           light_p1=light[0], light_ea=light[1], light_eb=light[2], **kw)

    image_mode = args.sharding == "image"
    full = None
    peers = []
    if image_mode:
        # every rank owns the full image; the other ranks' images are mapped as peer memory (CUDA IPC over NVLink)
        full = dev.buffer(npix * 16)
        full.clear()
        dev.sync()
        if world > 1:
            handles = [None] * world
            dist.all_gather_object(handles, full.ipc_export())
            peers = [dev.ipc_import(handles[r], npix * 16) for r in range(world) if r != rank]

    def step(i, profile_ctr=None):
        """one pass of the hot path over one frame, inputs resident in HBM"""
        if image_mode:  # this rank's tiles of frame(s) i, delivered into every rank's image by the resolve kernel
            return dev.render_gather(scene, params(i * args.spp_per_step, n_frames=args.spp_per_step, shard_index=rank,
                                                   shard_count=world, shard_block=64), full, peers, want_counters=profile_ctr)
        ctr = dev.render(scene, params(i * world + rank), frame, None, want_counters=profile_ctr)
        accum_t.add_(frame_t)
        return ctr

    # exact ray counts per step (untimed): counters force a sync, so they are read outside the timed region
    rays_per_step = []
    for i in range(args.steps):
        c = step(i, profile_ctr=True)
        rays_per_step.append(c["rays_closest"] + c["rays_any"])
    stat_ctr = dev.render(scene, params(rank, collect_stats=1, **({"shard_index": rank, "shard_count": world, "shard_block": 64} if image_mode else {})),
                          frame, None, want_counters=True)
    accum_t.zero_()

    sampler = ClockSampler(local)
    sampler.start()  # nvidia-smi needs ~100 ms to come up: start it before the warm-up, keep only timed-region rows
    for i in range(args.warmup):
        step(i)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()

    dev.profile(True)
    dev.profile_read()
    sampler.mark_begin()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps + 1)]
    for i in range(args.steps):
        flush.fill_(i & 255)           # evict L2 between timed iterations (untimed)
        ev[i][0].record(stream)
        step(i)
        ev[i][1].record(stream)
    # the job's single exchange: combine the per-rank accumulators (timed, once)
    ev[-1][0].record(stream)
    if world > 1 and not image_mode:
        dist.reduce(accum_t, dst=0, op=dist.ReduceOp.SUM)
    ev[-1][1].record(stream)
    torch.cuda.synchronize()
    sampler.mark_end()
    clocks = sampler.stop()
    prof = dev.profile_read()
    dev.profile(False)
    ms_steps = sum(a.elapsed_time(b) for a, b in ev[:-1])
    ms_exchange = ev[-1][0].elapsed_time(ev[-1][1])
    ms_total = ms_steps + ms_exchange
    t = torch.tensor([ms_total, float(sum(rays_per_step))], dtype=torch.float64, device="cuda")
    if world > 1:
        tmax = t.clone()
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        tsum = t.clone()
        dist.all_reduce(tsum, op=dist.ReduceOp.SUM)
        ms_total_max, rays_all = float(tmax[0]), float(tsum[1])
    else:
        ms_total_max, rays_all = ms_total, float(sum(rays_per_step))
    value = rays_all / (ms_total_max * 1e-3) / 1e6
    samples_all = npix * args.steps * (args.spp_per_step if image_mode else world)

    # ---- e2e: the same metric through the host-buffer C-ABI call (H2D scene records, D2H frame) ----------
    e2e = None
    if not args.no_e2e and not image_mode:
        dev2 = pt.Device(local)
        # page-locked host buffers: the caller's records (inputs) and two result frames (double buffering)
        tp, mp = pt.PinnedArray(tris.shape, tris.dtype), pt.PinnedArray(mats.shape, mats.dtype)
        tp.array[...] = tris
        mp.array[...] = mats
        outs = [pt.PinnedArray((npix, 4), np.float32), pt.PinnedArray((npix, 4), np.float32)]

        outs8 = [pt.PinnedArray((npix, 3), np.uint8), pt.PinnedArray((npix, 3), np.uint8)]

        def e2e_loop(pipelined, rgb8=False):
            rays = 0
            prev = None
            dst = outs8 if rgb8 else outs
            extra = {"output": pt.OUTPUT_RGB8} if rgb8 else {}
            t0 = time.perf_counter()
            for i in range(args.steps):
                job = dev2.render_host_async(tp.array, mp.array, params(i * world + rank, **extra), dst[i & 1].array)
                if not pipelined:
                    dev2.job_wait(job)
                else:
                    if prev is not None:
                        dev2.job_wait(prev)   # frame i-1 is now in host memory while frame i renders
                    prev = job
                rays += rays_per_step[i]
            if prev is not None:
                dev2.job_wait(prev)
            return time.perf_counter() - t0, rays

        for _ in range(2):
            e2e_loop(True)
        if world > 1:
            dist.barrier()
        t_sync, _ = e2e_loop(False)
        if world > 1:
            dist.barrier()
        t_e2e, e2e_rays = e2e_loop(True)
        checksum = float(outs[(args.steps - 1) & 1].array[:, 0].sum())  # the result is really on the host
        if world > 1:
            dist.barrier()
        e2e_loop(True, rgb8=True)
        if world > 1:
            dist.barrier()
        t_rgb8, _ = e2e_loop(True, rgb8=True)
        te = torch.tensor([t_e2e, float(e2e_rays), t_sync, t_rgb8], dtype=torch.float64, device="cuda")
        if world > 1:
            a_ = te.clone(); dist.all_reduce(a_, op=dist.ReduceOp.MAX)
            b_ = te.clone(); dist.all_reduce(b_, op=dist.ReduceOp.SUM)
            t_e2e, e2e_rays, t_sync, t_rgb8 = float(a_[0]), float(b_[1]), float(a_[2]), float(a_[3])
        e2e = {"value": e2e_rays / t_e2e / 1e6, "unit": "Mrays/s",
               "h2d_bytes_per_step": int(tris.nbytes + mats.nbytes), "d2h_bytes_per_step": int(npix * 16),
               "ms_per_step": t_e2e / args.steps * 1e3,
               "synchronous_value": e2e_rays / t_sync / 1e6, "synchronous_ms_per_step": t_sync / args.steps * 1e3,
               "host_checksum": checksum,
               "rgb8": {"value": e2e_rays / t_rgb8 / 1e6, "unit": "Mrays/s", "d2h_bytes_per_step": int(npix * 3), "ms_per_step": t_rgb8 / args.steps * 1e3,
                        "note": "same loop with params.output = RGB8: the reference's sqrt/x255/truncate output transform runs on the device "
                                "and 3 bytes per pixel travel instead of 16"},
               "timing": "host wall clock over K x {ptb_render_host_async (H2D records, render, D2H float4 frame into pinned host memory), "
                         "ptb_job_wait of the previous step}: D2H of step i overlaps the render of step i+1; synchronous_* waits every step. "
                         "Unlike the device-timed steps there is no L2 flush and no accumulate pass between e2e steps, so this can exceed `value`"}
        for pa in (tp, mp, *outs, *outs8):
            pa.free()
        dev2.close()

    # ---- roofline of the dominant kernel (the integrator launch) --------------------------------------------
    hbm_peak, peak_src, sm_max_mhz = peaks()
    n_rays_1 = rays_per_step[0]
    nodes_per_ray = stat_ctr["nodes"] / max(1, stat_ctr["rays_closest"] + stat_ctr["rays_any"])
    tests_per_ray = stat_ctr["tri_tests"] / max(1, stat_ctr["rays_closest"] + stat_ctr["rays_any"])
    q_bytes_per_ray = 16.0 * npix / n_rays_1  # one float4 radiance sample written per pixel-frame by the integrator
    node_bytes = {4: 128, 2: 64, 1: 32}[scene.info()["width"]]
    bytes_per_ray = nodes_per_ray * node_bytes + tests_per_ray * 48 + q_bytes_per_ray
    integ_ms = prof["integrator_ms"] / max(1, prof["batches"])
    rays_per_launch = n_rays_1 / max(1, prof["batches"] / args.steps)
    achieved_gbs = bytes_per_ray * rays_per_launch / (integ_ms * 1e-3) / 1e9
    sm_count = dev.sm_count()
    # fp32 lane-op estimate per ray: slab 24 per box, 2 boxes per node; MT stages from the oracle's exact stage counts
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    if os.path.exists(tpath):  # dram__bytes_read.sum + dram__bytes_write.sum per launch, from the committed ncu --set full capture
        traffic = json.load(open(tpath)).get(args.workload, {}).get("dram_bytes_per_launch")
    roof = {"bound": "hbm", "achieved": achieved_gbs, "peak": hbm_peak, "unit": "GB/s", "frac": achieved_gbs / hbm_peak,
            "traffic": traffic, "peak_source": peak_src, "kernel": "k_mega<AO>" if wl["mode"] == 1 else "integrator",
            "kernel_ms": integ_ms, "kernel_share_of_step": prof["integrator_ms"] / ms_steps if ms_steps else None,
            "bytes_per_ray": bytes_per_ray, "nodes_per_ray": nodes_per_ray, "tri_tests_per_ray": tests_per_ray,
            "note": "algorithmic bytes = nodes*(128 B 4-wide | 64 B binary) + tri_tests*48 + 16 B/sample (SURVEY 8d); the 5 KB scene is shared-memory resident, so this "
                    "is the ON-CHIP stream, not DRAM traffic: the binding limit is fp32 issue (see simt)"}

    # ---- CPU baseline beside it (rank 0, N=1 only) -------------------------------------------------------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu = cpu_reference(wl, seconds_budget=12.0, sample_px=None if not wl.get("tess") else 16 * 16, max_frames=64)
        one = cpu_reference(wl, seconds_budget=3.0, sample_px=256 * 256 if not wl.get("tess") else 8 * 8, max_frames=16, n_threads=1)
        cpu["single_thread"] = {"value": one["value"], "unit": "Mrays/s", "cores": 1, "sample": one["sample"]}
        ref_cmp = os.path.join(ROOT, "profiles", "reference_opencl_r01", "compare.json")
        if os.path.exists(ref_cmp):  # recorded once, not re-measured here: the unmodified reference on this GPU via OpenCL
            cpu["reference_opencl_on_b200"] = {"raycast_10000_frames_512x512_s": 357.4, "same_flow_through_libptb200_s": 5.0,
                                               "image_rrmse": json.load(open(ref_cmp))["rrmse"],
                                               "source": "profiles/reference_opencl_r01 (tools/run_reference_opencl.sh)"}
        from oracle import binding as ob
        # stage counts of the BVH path at reduced size -> fp32 lane-ops per ray (SURVEY 8d formula)
        otris, omats = ob.load_model(SCENE)
        if not wl.get("tess"):
            b = ob.build_bvh(otris, width=scene.info()["width"])
            bvh, _keep = ob.make_bvh(b["nodes"], b["tri_order"])
            op = ob.default_params(256, 256, n_frames=1, mode=wl["mode"], accum=1, use_bvh=1, ao_samples=wl.get("ao_samples", 16),
                                   max_depth=wl.get("max_depth", 16), light_p1=light[0], light_ea=light[1], light_eb=light[2])
            _, _, oc = ob.render(op, otris, omats, bvh=bvh)
            nr = oc["rays_closest"] + oc["rays_any"]
            boxes = oc["nodes"] * (node_bytes // 32) if scene.info()["width"] != 1 else nr * scene.info()["n_nodes"]  # FLAT: every leaf box, every ray
            flops_per_ray = (boxes * 24 + oc["tri_tests"] * 14 + oc["tri_u"] * 10 + oc["tri_v"] * 15 + oc["tri_t"] * 6 + oc["tri_accept"] * 40) / nr
            sm_mhz = clocks.get("sm_mhz") or sm_max_mhz
            peak_lane_ops = sm_count * 128 * sm_max_mhz * 1e6  # no FMA in the parity build: 1 lane-op per lane per clock
            ach = flops_per_ray * (rays_per_launch / (integ_ms * 1e-3))
            roof["simt"] = {"bound": "fp32_issue", "flops_per_ray": flops_per_ray, "achieved_tlaneops": ach / 1e12,
                            "peak_tlaneops": peak_lane_ops / 1e12, "frac": ach / peak_lane_ops, "sm_mhz_during_run": sm_mhz,
                            "note": "algorithmic fp32 ops of traversal only (slab 24/box x boxes per node, MT stages 14/10/15/6/40); shading, RNG and sincos are extra"}
        cpu.pop("seconds"); cpu.pop("rays"); cpu.pop("frames"); cpu.pop("wh")

    ab = None
    if args.ab and world == 1:
        ab = {}
        variants = [("megakernel_bvh", dict(integrator=pt.INTEGRATOR_MEGAKERNEL, accel=pt.ACCEL_BVH)),
                    ("wavefront_bvh", dict(integrator=pt.INTEGRATOR_WAVEFRONT, accel=pt.ACCEL_BVH))]
        if len(tris) <= 4096:  # the reference's brute-force loop is O(triangles) per ray
            variants.append(("megakernel_brute", dict(integrator=pt.INTEGRATOR_MEGAKERNEL, accel=pt.ACCEL_BRUTE)))
        for name, kw in variants:
            def one(i):
                p = make_params(pt, wl, first_frame=i, light_p1=light[0], light_ea=light[1], light_eb=light[2], **kw)
                dev.render(scene, p, frame, None)
            for i in range(3):
                one(i)
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            tot = 0.0
            for i in range(args.steps):
                flush.fill_(i & 255)
                a.record(stream); one(i); b.record(stream)
                torch.cuda.synchronize()
                tot += a.elapsed_time(b)
            ab[name] = {"Mrays/s": sum(rays_per_step) / (tot * 1e-3) / 1e6, "ms_per_step": tot / args.steps}

    if rank == 0:
        line = {
            "metric": "Mrays/sec (primary+secondary)", "value": value, "unit": "Mrays/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_total_max / args.steps,
            "higher_is_better": True, "scaling": "strong" if image_mode else "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"{args.workload}: {wl['desc']}", "scene_triangles": int(len(tris)),
                       "rays_per_step_per_gpu": int(n_rays_1),
                       "step": (f"{args.spp_per_step} frame(s) of ONE image tile-sharded over the ranks (64-pixel blocks round-robin)" if image_mode
                                else "one frame (1 spp) of the full image per rank; rank r renders frame step*N+r"),
                       "integrator": args.integrator, "accel": args.accel, "l2": "256 MiB flush between timed steps", "cpu_affinity": affinity,
                       "parallelism": (f"image tiles over {world} GPU(s); the gather is fused into the resolve kernel (stores into every rank's image "
                                       f"over NVLink peer memory), no collective" if image_mode
                                       else f"frames dealt round-robin over {world} GPU(s); one NCCL reduce of the accumulators at the end")},
            "spp_per_s": args.steps * (args.spp_per_step if image_mode else world) / (ms_total_max * 1e-3), "msamples_per_s": samples_all / (ms_total_max * 1e-3) / 1e6,
            "exchange_ms": ms_exchange, "bvh_build_s": build_s, "sm_count": sm_count,
            "clocks": clocks, "e2e": e2e, "gpu_launches": int(prof["kernel_launches"]),
            "roofline": roof, "cpu_baseline": cpu,
        }
        if ab:
            line["ab"] = ab
        print(json.dumps(line), flush=True)
    if image_mode:
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        for b in peers:
            b.close()
        if world > 1:
            dist.barrier()
        full.close()
    frame.close()
    scene.close()
    dev.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
