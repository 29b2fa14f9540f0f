#!/usr/bin/env python
"""bench.py — throughput of the hot path (Renderer::Accumulate x spp + Renderer::Render) on N B200s of one node.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload c1|c2|c3|c4] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P bench.py --gpus N ...

One STEP = one frame of the workload: ResetAccumulator, Accumulate() for the workload's sample count, Render().
Default workload = BASELINE.json configs[1] (C2): Scenes::Default, 1920x1080 (rendered 1920x1088), 64 spp,
median-of-means with 8 buckets, max_bounces 16, MIS on. `value` is Mrays/s (extension + shadow rays actually traced, counted on
the device) with the scene resident in HBM; `e2e` is the same through the public API with host buffers (scene upload from host,
frame download to pinned host memory every step). Prints ONE JSON line on rank 0.

--impl reference times the reference's CPU path. The reference itself cannot be built here (MSVC-only C++, glm/VCL/PPL absent,
SURVEY §8c), so this arm runs the oracle port (oracle/liboracle_fast.so, all host threads) on a bounded sample of the same
workload. It is the ONE place besides cpu_baseline where bench.py executes oracle/ code, as the thing being measured on the CPU.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.join(ROOT, "cpu-raytracing-experiments_b200")
sys.path[:0] = [ROOT, PKG]

WORKLOADS = {
    # name: scene, width, height(internal, multiple of 16), spp, max_bounces, K
    "c1": dict(desc="C1 default 9-sphere scene 1280x720 1spp max_bounces=8 MIS K=5 (1 sample lands in bucket 1)", scene="default", w=1280, h=720, spp=5, mb=8, K=5),
    "c2": dict(desc="C2 default 9-sphere scene 1920x1080 (internal 1920x1088) 64spp median-of-means K=8 max_bounces=16 MIS brute-force", scene="default", w=1920, h=1088, spp=64, mb=16, K=8),
    "c3": dict(desc="C3 random 100k-sphere BVH scene 1920x1080 (internal 1920x1088) 16spp K=8 max_bounces=16 NEE+MIS", scene="random100000", w=1920, h=1088, spp=16, mb=16, K=8),
    "c4": dict(desc="C4 random 1M-sphere BVH scene 3840x2160 32spp K=8 max_bounces=16", scene="random1000000", w=3840, h=2160, spp=32, mb=16, K=8),
    # C5: progressive convergence run; the frame's 1024 samples are SPLIT over the ranks by bucket (strong scaling)
    "c5": dict(desc="C5 progressive 3840x2160 1024spp on the C3 100k-sphere scene, K=8 buckets split over the GPUs, max_bounces=16", scene="random100000", w=3840, h=2160, spp=1024, mb=16, K=8, strong=True),
}
METRIC, UNIT = "Mrays/s", "Mrays/s"


def make_scene(name):
    import scenes
    if name == "default":
        return scenes.default_scene()
    return scenes.random_scene(int(name[len("random"):]))


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p)); return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)", float(d.get("sm_max_mhz", 1965.0))
    return 6650.0, "fallback (B200_PROFILING.md)", 1965.0


class ClockSampler:
    """SM clock + throttle reasons sampled DURING the timed region (B200_PROFILING.md): NVML polled every 5 ms from a thread
    (nvidia_ml_py), falling back to `nvidia-smi -lms` when NVML cannot be loaded."""
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, gpu):
        self.gpu = gpu; self.sm = []; self.reasons = set(); self.mx = None; self.power = []; self.t = None; self.p = None; self.f = None; self.run = False

    def start(self):
        try:
            import threading
            import pynvml as nv
            nv.nvmlInit()
            idx = self.gpu
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            if vis:
                try:
                    idx = int(vis.split(",")[self.gpu])
                except Exception:
                    idx = self.gpu
            h = nv.nvmlDeviceGetHandleByIndex(idx)
            self.mx = float(nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM))
            bits = {"hw_slowdown": 0x8, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20, "sw_power_cap": 0x4}
            get_reasons = getattr(nv, "nvmlDeviceGetCurrentClocksEventReasons", None) or nv.nvmlDeviceGetCurrentClocksThrottleReasons
            self.run = True

            def loop():
                while self.run:
                    try:
                        self.sm.append(float(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)))
                        r = get_reasons(h)
                        for k, b in bits.items():
                            if r & b:
                                self.reasons.add(k)
                        self.power.append(nv.nvmlDeviceGetPowerUsage(h) / 1e3)
                    except Exception:
                        pass
                    time.sleep(0.005)
            self.t = threading.Thread(target=loop, daemon=True); self.t.start()
        except Exception:
            self.t = None
            try:
                self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
                self.p = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "20"], stdout=self.f, stderr=subprocess.DEVNULL)
                time.sleep(0.3)
            except Exception:
                self.p = None

    def stop(self):
        if self.t is not None:
            self.run = False; self.t.join(timeout=1.0)
            return {"sm_mhz": statistics.median(self.sm) if self.sm else None, "sm_max_mhz": self.mx, "reasons": sorted(self.reasons), "samples": len(self.sm),
                    "power_w_max": max(self.power) if self.power else None, "source": "nvml"}
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["clock sampling unavailable"]}
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush(); self.f.seek(0)
        sm, mx, reasons = [], [], set()
        for line in self.f.read().strip().splitlines():
            c = [x.strip() for x in line.split(",")]
            if len(c) < 7:
                continue
            try:
                sm.append(float(c[0])); mx.append(float(c[1]))
            except ValueError:
                continue
            for name, v in zip(["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"], c[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        os.unlink(self.f.name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons), "samples": len(sm), "source": "nvidia-smi"}


# bounded CPU samples: full frames for the 9-sphere scene; a fixed random subset of 16x16 tiles for the BVH scenes (every tile is an
# independent unit of the reference's parallel_for, Renderer.hpp:75-84), sized for roughly 5-20 s of host time
CPU_TILE_SAMPLE = {"c3": 1024, "c4": 384, "c5": 1024}


def cpu_reference_run(wl, scene, samples, first_sample=0):
    """Times THE REFERENCE ITSELF — Renderer<>::Accumulate from /root/reference's Renderer.hpp, compiled into oracle/_ref/librefrenderer.so
    by oracle/ref_renderer_build.sh (brute force, as shipped: USEBVH false) — on all host threads. Rays are not counted by the reference;
    they are counted by the oracle's slot-exact mode on the same samples (bit-identical paths, tests/test_oracle_ref_renderer.py), untimed."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import oracle_py
    threads = os.cpu_count() or 1
    r = oracle_py.ReferenceRenderer(scene, wl["w"], wl["h"], wl["mb"])
    r.set_accumulations(first_sample)
    t0 = time.perf_counter()
    r.accumulate(samples)
    dt = time.perf_counter() - t0
    r.close()
    o = oracle_py.Oracle(wl["w"], wl["h"], max_bounces=wl["mb"], K=5, flags=oracle_py.ORC_SLOT_EXACT, fast=True)
    o.set_scene(scene); o.set_accumulations(first_sample); o.reset_counters(); o.accumulate(samples, threads=threads)
    c = o.counters(); rays = c["extension_rays"] + c["shadow_rays"]; o.close()
    return dict(seconds=dt, rays=rays, paths=wl["w"] * wl["h"] * samples, threads=threads, what=f"{samples} spp of the {wl['w']}x{wl['h']} frame",
                mode="the reference's own Renderer::Accumulate (Renderer.hpp, g++ -O2 -mavx2 -mfma, brute force as shipped), tiles over all host threads",
                kind="reference")


def cpu_oracle_run(wl, scene, samples, fast=True, workload_name=None):
    """Times the oracle port on all host threads: `samples` x Accumulate at the workload's size (+ Render when a round completes)."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import numpy as np
    import oracle_py
    threads = os.cpu_count() or 1
    bvh = wl["scene"] != "default"
    o = oracle_py.Oracle(wl["w"], wl["h"], max_bounces=wl["mb"], K=wl["K"], flags=oracle_py.ORC_BVH if bvh else 0, fast=fast)
    o.set_scene(scene)
    o.reset_counters()
    n_tiles_all = (wl["w"] // 16) * (wl["h"] // 16)
    n_tiles = min(n_tiles_all, CPU_TILE_SAMPLE.get(workload_name, n_tiles_all)) if bvh else n_tiles_all
    t0 = time.perf_counter()
    if n_tiles < n_tiles_all:
        tiles = np.random.RandomState(1).choice(n_tiles_all, n_tiles, replace=False).astype(np.uint32)
        o.accumulate_tiles(tiles, samples, threads=threads)
    else:
        o.accumulate(samples, threads=threads); o.render()
    dt = time.perf_counter() - t0
    c = o.counters(); rays = c["extension_rays"] + c["shadow_rays"]
    o.close()
    what = f"{samples} spp of {n_tiles} of the {n_tiles_all} 16x16 tiles of the {wl['w']}x{wl['h']} frame" if n_tiles < n_tiles_all else f"{samples} spp of the {wl['w']}x{wl['h']} frame"
    return dict(seconds=dt, rays=rays, paths=n_tiles * 256 * samples, threads=threads, what=what,
                mode="stream-BVH (BVH.hpp:320-358 restated)" if bvh else "brute force (as shipped, USEBVH false)")


def run_reference(args, wl, rank):
    """--impl reference: the CPU path on the box's host cores, one bounded sample per step."""
    if rank != 0:
        return
    scene = make_scene(wl["scene"])
    per_step = 4 if wl["scene"] == "default" else 1  # a few spp of the workload's frame per step: a bounded sample (~0.1-20 s on the host cores)
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import oracle_py
    # The reference itself when it is runnable for this workload: it ships brute force only (USEBVH false), so the 9-sphere
    # configurations; max_bounces must be one of the instantiated template values. The BVH workloads keep the oracle port
    # (stream-BVH restatement): brute force over 1e5-1e6 spheres is ~1e11-1e12 sphere tests per bounce pass.
    use_ref = wl["scene"] == "default" and oracle_py.have_reference_renderer() and wl["mb"] in oracle_py.REF_MAX_BOUNCES
    run = (lambda k: cpu_reference_run(wl, scene, per_step, first_sample=k * per_step)) if use_ref else (lambda k: cpu_oracle_run(wl, scene, per_step, workload_name=args.workload))
    for k in range(args.warmup):
        run(k)
    secs, rays, paths = 0.0, 0, 0
    threads = mode = None
    for k in range(args.steps):
        r = run(k); secs += r["seconds"]; rays += r["rays"]; paths += r["paths"]; threads, mode = r["threads"], r["mode"]; what = r["what"]
    v = rays / secs / 1e6
    kind = "reference" if use_ref else "port"
    sample = f"{what} per step ({paths // args.steps} paths), {args.steps} steps; " + (mode if use_ref else f"oracle port -O3 -march=native, {mode}")
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * secs / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": wl["desc"]}, "paths_per_s": paths / secs,
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": threads, "kind": kind, "sample": sample},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "note": ("oracle/_ref/librefrenderer.so: the reference's Renderer.hpp and the headers it includes, compiled from /root/reference by oracle/ref_renderer_build.sh "
                 "(stand-ins for ppl.h / Image.h / glm / VCL, six token-level syntax edits)") if use_ref else
                "the reference ships brute force only; for BVH workloads (or without oracle/_ref) the oracle port is timed",
    }))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1); ap.add_argument("--steps", type=int, default=20); ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"]); ap.add_argument("--workload", default="c2", choices=sorted(WORKLOADS))
    ap.add_argument("--no-cpu-baseline", action="store_true"); ap.add_argument("--no-profile-pass", action="store_true")
    ap.add_argument("--samples-in-flight", type=int, default=0, help="0 = library default (auto)")
    ap.add_argument("--no-graph", action="store_true", help="launch kernel by kernel (for ncu); never used for a reported number")
    ap.add_argument("--reference-exact", action="store_true", help="brute-force workloads: B2R_FLAG_REFERENCE_EXACT (bit-identical to the reference's own renderer; one extra ranking kernel per bounce)")
    ap.add_argument("--combine", default="p2p", choices=["p2p", "nccl"], help="multi-GPU frame combine: resolve kernel reads peer buckets over NVLink (p2p) or NCCL all-reduce then resolve")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 0)
    wl = WORKLOADS[args.workload]
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, wl, rank); return

    import numpy as np
    import torch
    import torch.distributed as dist
    import __graft_entry__ as entry
    if rank == 0:
        entry.build()
    if world > 1:
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
        dist.barrier()
    import b2r, b2r_dist
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: libb2r has no CPU fallback")
    dev = torch.device(f"cuda:{local}"); torch.cuda.set_device(dev)
    warm = max(args.warmup, 3)

    scene = make_scene(wl["scene"])
    ps = b2r.PreparedScene(scene, wl["w"], wl["h"])  # host BVH build + light list: scene (re)build, outside the hot path
    stream = torch.cuda.Stream(device=dev)  # a real (non-default) stream: the library launches on it and the timing events are recorded on it
    torch.cuda.set_stream(stream)
    K, spp = wl["K"], wl["spp"]
    base_flags = b2r.FLAG_REFERENCE_EXACT if (args.reference_exact and wl["scene"] == "default") else 0
    shard = b2r_dist.shard_kwargs(rank, world, K)
    r = b2r.Renderer(ps, wl["w"], wl["h"], max_bounces=wl["mb"], buckets=K, device=local, stream=stream.cuda_stream,
                     samples_in_flight=args.samples_in_flight, flags=(b2r.FLAG_NO_GRAPH if args.no_graph else 0) | base_flags, **shard)
    # weak scaling: every rank renders `spp` samples of its own buckets per step => world*spp sample indices per step
    # (C5 is the strong-scaling run: the frame's spp are divided among the ranks)
    strong = bool(wl.get("strong"))
    step_samples = spp if strong else spp * world
    local_buckets = b2r_dist.buckets_tensor(r, dev)
    use_p2p = world > 1 and args.combine == "p2p"
    if use_p2p:
        b2r_dist.open_peers(r)

    dbg = bool(os.environ.get("B2R_BENCH_DEBUG"))

    frame_no = [0]

    def step(to_host_fb=None, upload=False, async_fbs=None):
        t0 = time.perf_counter()
        if upload:
            r.SetScene(ps)  # host -> device: spheres, materials, lights, flattened BVH, camera
        t1 = time.perf_counter()
        r.ResetAccumulator()
        r.Accumulate(step_samples)
        if dbg:
            r.sync(); print(f"[dbg] upload {t1 - t0:.4f}s accumulate {time.perf_counter() - t1:.4f}s", file=sys.stderr)
        t2 = time.perf_counter()
        if use_p2p:
            # fused combine: rank 0's resolve kernel pulls every bucket from its owner over NVLink. The two barriers order it after
            # every rank's last bounce and before any rank's next ResetAccumulator.
            r.sync(); dist.barrier()
            ok = r.RenderPeers(to_host=to_host_fb is not None, out=to_host_fb) if rank == 0 else True
            dist.barrier()
        elif world > 1:
            combined = b2r_dist.combine_buckets(local_buckets)  # the one collective: NCCL all-reduce of the bucket sums
            ok = r.Render(to_host=to_host_fb is not None and rank == 0, out=to_host_fb, dev_buckets=combined.data_ptr())
        elif async_fbs is not None:
            # progressive rendering as a user would drive it: frame N is copied out on the library's second stream into one of two pinned
            # buffers while the samples of frame N+1 are traced (b2r_resolve_async); nothing is skipped, every frame lands on the host
            ok = r.RenderAsync(async_fbs[frame_no[0] & 1]); frame_no[0] += 1
        else:
            ok = r.Render(to_host=to_host_fb is not None, out=to_host_fb)
        assert ok
        if dbg:
            print(f"[dbg] render {time.perf_counter() - t2:.4f}s total {time.perf_counter() - t0:.4f}s", file=sys.stderr)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def timed(n, **kw):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(n):
            step(**kw)
        if kw.get("async_fbs") is not None:
            r.WaitFrame()  # the last frame has landed in host memory before the clock stops
        e1.record(stream)
        barrier()
        ms = e0.elapsed_time(e1)
        if dbg:
            print(f"[dbg] timed region {ms:.2f} ms for {n} steps", file=sys.stderr)
        if world > 1:
            t = torch.tensor([ms], device=dev); dist.all_reduce(t, op=dist.ReduceOp.MAX); ms = float(t.item())
        return ms

    for _ in range(warm):
        step()
    barrier()
    r.reset_counters()
    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()
    ms_total = timed(args.steps)
    clk = clocks.stop() if rank == 0 else None
    cnt = r.counters()
    rays_local = cnt["extension_rays"] + cnt["shadow_rays"]
    if world > 1:
        t = torch.tensor([rays_local, cnt["launches"]], device=dev, dtype=torch.float64); dist.all_reduce(t); rays_total, launches = float(t[0]), int(t[1])
    else:
        rays_total, launches = float(rays_local), cnt["launches"]
    secs = ms_total / 1e3
    value = rays_total / secs / 1e6
    paths_total = wl["w"] * wl["h"] * step_samples * args.steps

    # ---- e2e: same steps through the public API with HOST buffers (scene upload + frame download each step)
    fb_host = torch.empty((wl["h"], wl["w"], 4), dtype=torch.float32, pin_memory=True).numpy()
    e2e_kw = dict(to_host_fb=fb_host, upload=True)
    if world == 1:  # single GPU: frames leave through b2r_resolve_async into two alternating pinned buffers (multi-GPU: the fused P2P resolve stays synchronous)
        fb_host2 = torch.empty((wl["h"], wl["w"], 4), dtype=torch.float32, pin_memory=True).numpy()
        e2e_kw = dict(async_fbs=[fb_host, fb_host2], upload=True)
    for _ in range(2):
        step(**e2e_kw)
    if world == 1:
        r.WaitFrame()
    r.reset_counters()
    ms_e2e = timed(args.steps, **e2e_kw)
    c2 = r.counters(); rays_e2e = c2["extension_rays"] + c2["shadow_rays"]
    if world > 1:
        t = torch.tensor([rays_e2e], device=dev, dtype=torch.float64); dist.all_reduce(t); rays_e2e = float(t[0])
    wide, _ms = r.wide_nodes()
    h2d = len(ps.prims) * 20 + len(ps.material) * 32 + max(1, len(ps.lights)) * 32 + wide.shape[0] * 128 + 44
    d2h = wl["w"] * wl["h"] * 16 if rank == 0 else 0
    e2e_value = rays_e2e / (ms_e2e / 1e3) / 1e6

    # ---- per-kernel pass: the same K steps launched kernel by kernel with an event pair around every launch (roofline)
    roof = None; kernel_ms = None
    hbm_peak, peak_src, sm_max = peaks()
    if not args.no_profile_pass:
        r.set_flags(b2r.FLAG_NO_GRAPH | base_flags); r.SetCamera(ps.camera)
        step(); r.sync(); r.kernel_times(reset=True); r.reset_counters()
        for _ in range(args.steps):
            step()
        r.sync()
        kt = r.kernel_times(reset=True); pc = r.counters()
        r.set_flags(base_flags); r.SetCamera(ps.camera)
        kernel_ms = {k: {"ms": v[0], "launches": int(v[1])} for k, v in kt.items() if v[1]}
        total_kernel_ms = sum(v[0] for v in kt.values())
        ext, shadow, hits, events, dropped = pc["extension_rays"], pc["shadow_rays"], pc["shaded_hits"], pc["radiance_events"], pc["dropped"]
        primaries = wl["w"] * wl["h"] * (step_samples // world) * args.steps
        if kt["bounce_brute"][1]:
            # DESIGN.md "Algorithmic bytes": 44 B path record read per non-primary ray + 44 B written per continuing path (= every
            # non-primary ray was written once) + 12 B radiance entry started per primary path + 24 B radiance read-modify-write
            # per contribution + 12 B per dropped path
            name, ms, n = "k_bounce_brute", kt["bounce_brute"][0], kt["bounce_brute"][1]
            alg = 88.0 * (ext - primaries) + 12.0 * primaries + 24.0 * events + 12.0 * dropped
        else:
            # dominant kernel of the BVH pipeline = closest-hit traversal: 32 B ray read + 8 B hit written per extension ray from HBM;
            # node/sphere fetches are L2-resident and reported separately through the box/sphere counters
            name, ms, n = "k_intersect_closest", kt["intersect_closest"][0], kt["intersect_closest"][1]
            alg = 40.0 * ext
        achieved = alg / (ms / 1e3) / 1e9
        roof = {"bound": "hbm", "kernel": name, "achieved": achieved, "peak": hbm_peak, "unit": "GB/s", "frac": achieved / hbm_peak, "traffic": None,
                "peak_source": peak_src, "avg_launch_ms": ms / n, "launches": int(n), "algorithmic_bytes_per_launch": alg / n,
                "kernel_share_of_step": ms / total_kernel_ms if total_kernel_ms else None,
                "timing": "CUDA event pair around every launch, non-graph pass of the same steps on the launching stream"}
        # FP32 view (never tensor cores): 20 flop per sphere test (BVH.hpp:251-265), ~250 per shaded hit
        if kt["bounce_brute"][1]:
            n_prims = len(ps.prims)
            flops = 20.0 * n_prims * (ext + shadow) + 250.0 * hits  # shadow: upper bound (any-hit exits early)
            fp32_peak = 148 * 128 * 2 * sm_max * 1e6 / 1e12
            roof["fp32"] = {"achieved_tflops": flops / (ms / 1e3) / 1e12, "peak_tflops": fp32_peak, "frac": flops / (ms / 1e3) / 1e12 / fp32_peak,
                            "note": "accounting flops (20/sphere test, 250/shaded hit); shadow tests counted at their upper bound"}
        ncu = os.path.join(ROOT, "profiles", "r01_traffic.json")
        if os.path.exists(ncu):
            try:
                tr = json.load(open(ncu)).get(args.workload, {}).get(name)
                if tr:
                    roof["traffic"] = tr["dram_bytes_per_launch"]; roof["traffic_source"] = tr.get("source")
                    if tr.get("issue_slots_busy_pct") is not None:  # the binding resource of these kernels (ncu, not live): instruction issue
                        roof["issue"] = {"slots_busy_pct": tr["issue_slots_busy_pct"], "source": "ncu smsp__issue_active, same capture as traffic"}
            except Exception:
                pass

    # ---- the same workload in B2R_FLAG_REFERENCE_EXACT mode (bit-identical to the reference's own renderer), reported beside the
    # default mode's number: brute-force workloads, N=1 only, same steps / warm-up / timing as the headline value
    exact_line = None
    if world == 1 and not base_flags and wl["scene"] == "default" and not args.no_profile_pass:
        rx = b2r.Renderer(ps, wl["w"], wl["h"], max_bounces=wl["mb"], buckets=K, device=local, stream=stream.cuda_stream,
                          samples_in_flight=args.samples_in_flight, flags=b2r.FLAG_REFERENCE_EXACT)
        def step_x():
            rx.ResetAccumulator(); rx.Accumulate(step_samples); assert rx.Render(to_host=False)
        for _ in range(warm):
            step_x()
        torch.cuda.synchronize(dev); rx.reset_counters()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(args.steps):
            step_x()
        e1.record(stream); torch.cuda.synchronize(dev)
        ms_x = e0.elapsed_time(e1); cx = rx.counters(); rays_x = cx["extension_rays"] + cx["shadow_rays"]
        exact_line = {"value": rays_x / (ms_x / 1e3) / 1e6, "unit": UNIT, "ms_per_step": ms_x / args.steps,
                      "note": "B2R_FLAG_REFERENCE_EXACT: the reference's per-tile stream order and scalar-tail sphere formula; results bit-identical to the reference's own Renderer::Accumulate/Render (tests/test_gpu_parity.py)"}
        rx.close()

    # ---- scene edit (the reference rebuilds its BVH on every geometry drag, Application.cpp:508-510): every sphere of the
    # BVH workload moves by up to a quarter of its radius; b2r_refit_scene (GPU refit of the traversal tree, topology kept) beside the
    # full rebuild (host reference-BVH build + b2r_upload_scene: traversal-tree build, flatten, H2D). Outside the timed region; N=1 only.
    edit_line = None
    if rank == 0 and world == 1 and wl["scene"] != "default" and not args.no_profile_pass:
        rs = np.random.RandomState(5)
        geo2 = ps.geometry.copy(); rad = np.sqrt(geo2["radius_sq"])
        geo2["position"] += (rs.uniform(-1, 1, (len(geo2), 3)) * (0.25 * rad)[:, None]).astype(np.float32)

        def steps_ms(n=3):
            step(); torch.cuda.synchronize(dev)
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(stream)
            for _ in range(n):
                step()
            b.record(stream); torch.cuda.synchronize(dev)
            return a.elapsed_time(b) / n
        r.sync(); t0 = time.perf_counter()
        r.RefitScene(geo2, want_quality=False, keep_order=True); r.sync()
        refit_ms = (time.perf_counter() - t0) * 1e3
        t0 = time.perf_counter()
        r.RefitScene(geo2, want_quality=False); r.sync()   # as the app does it: reference BVH rebuilt on the host (new leaf order), leaves re-linked
        refit_reordered_ms = (time.perf_counter() - t0) * 1e3
        q = r.RefitScene(geo2, keep_order=True)
        ms_refit = steps_ms()
        t0 = time.perf_counter()
        scene2 = dict(scene); scene2["geometry"] = geo2
        ps2 = b2r.PreparedScene(scene2, wl["w"], wl["h"])
        t1 = time.perf_counter()
        r.SetScene(ps2); r.sync()
        t2 = time.perf_counter()
        ms_rebuilt = steps_ms()
        edit_line = {"moved": "every sphere by up to 0.25 radius", "refit_ms": refit_ms, "refit_with_host_reference_bvh_rebuild_ms": refit_reordered_ms, "quality_ratio": q,
                     "rebuild_ms": {"reference_bvh_host": (t1 - t0) * 1e3, "upload_scene": (t2 - t1) * 1e3},
                     "ms_per_step_after_refit": ms_refit, "ms_per_step_after_rebuild": ms_rebuilt,
                     "note": "refit = match prims to geometry + pack + H2D of the spheres and lights + k_refit_level per tree level, host wall clock incl. stream sync; leaf order kept (refit_ms) or the reference BVH rebuilt on the host first, as Application.cpp:508 does"}
        r.SetScene(ps)

    # ---- CPU baseline beside it (rank 0, N=1 only): the oracle port on a bounded sample of the same workload, on the box's host
    # cores. Run as a fresh `bench.py --impl reference` process so that its threads see the same conditions as the reference arm
    # the driver launches (inside this process, after CUDA/torch start-up, the same code ran at about half the speed).
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        try:
            out = subprocess.run([sys.executable, os.path.abspath(__file__), "--impl", "reference", "--workload", args.workload, "--steps", "2", "--warmup", "1"],
                                 capture_output=True, text=True, timeout=900).stdout.strip().splitlines()[-1]
            ref = json.loads(out)
            cpu = dict(ref["cpu_baseline"]); cpu["paths_per_s"] = ref.get("paths_per_s")
        except Exception as e:  # never let the baseline leg break the GPU line
            cpu = {"value": None, "unit": UNIT, "cores": os.cpu_count(), "kind": "port", "sample": f"failed: {e}"}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": warm, "ms_per_step": ms_total / args.steps,
            "higher_is_better": True, "scaling": "strong" if strong else "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": wl["desc"] + (" [B2R_FLAG_REFERENCE_EXACT: reference's stream order and scalar-tail formula]" if base_flags else ""), "samples_per_step_per_gpu": step_samples // world, "samples_in_flight": args.samples_in_flight,
                       "partition": (f"sample buckets b%{world}==rank, scene+BVH replicated; " + ("resolve kernel on rank 0 reads peer bucket arrays over NVLink (CUDA IPC), 2 barriers per frame" if use_p2p else "one NCCL all-reduce of bucket sums per frame")) if world > 1 else "single GPU",
                       "l2": "no flush needed: each step streams >1 GB of path-queue records (>> 126 MB L2); RNG-unique samples every step"},
            "paths_per_s": paths_total / secs, "rays_per_step": rays_total / args.steps,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h), "ms_per_step": ms_e2e / args.steps,
                    "includes": "upload_scene (pack + 128-B BVH flatten on host, H2D), set_camera, reset, Accumulate x spp, Render, D2H of the RGBA32F frame into pinned memory"
                                + ("; frames leave through b2r_resolve_async (copy of frame N overlaps the tracing of frame N+1, two pinned buffers, last frame waited for inside the timed region)" if world == 1 else "")},
            "gpu_launches": int(launches), "clocks": clk, "roofline": roof, "cpu_baseline": cpu, "kernel_ms": kernel_ms,
        }
        if exact_line:
            line["reference_exact_mode"] = exact_line
        if edit_line:
            line["scene_edit"] = edit_line
        print(json.dumps(line))
    r.close()
    if world > 1:
        dist.barrier(); dist.destroy_process_group()


if __name__ == "__main__":
    main()
