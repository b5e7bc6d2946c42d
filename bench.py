#!/usr/bin/env python
"""Benchmark of the MultiSPA geometry hot path (BASELINE.json metric: annotated frame-pair samples/s).

    python bench.py [--gpus N] [--steps K] [--warmup W]            # this repo's CUDA path
    python bench.py --impl reference [...]                           # the reference's own CPU code (oracle/_ref)

Workload at N=1 = BASELINE.json configs[1]: "ScanNet visual_correspondence: 100 scenes x 1k frame-pairs,
~200k pts/scene, 1xB200", independent-pair model of SURVEY.md §8(d): every pair owns its two frames, so a scene
carries 2000 frames (poses + 640x480 u16 depth maps) and a step projects 100 x 2000 frames and annotates 100 x 1000
pairs.  A "step" is one pass over all scenes resident in HBM; with N GPUs every rank owns its own scenes (weak
scaling, no collective on the data path; one host-side gather of the per-rank manifest sizes).

One JSON line is printed by rank 0.  See DESIGN.md §Measurement for how each field is obtained.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "annotated_frame_pairs_per_sec"
UNIT = "pairs/s"
DEPTH_BYTES = 640 * 480 * 2
TILE_POINTS = 1024            # msp_cull9.cuh: kC9TilePoints
TILE_PREP_BYTES = 13904       # sizeof(CullTilePrep)


def algorithmic_bytes_frame(n: int) -> float:
    """B_frame(N) = 12N (xyz f32) + 614400 (depth u16) + 80 (pose, intrinsics) + N/8 (mask out)."""
    return 12.0 * n + DEPTH_BYTES + 80.0 + n / 8.0


def compulsory_bytes_launch(n: int, frames: int, stride_words: int, n_objects: int) -> float:
    """HBM bytes one K1 call has to move WITH the kernel's own reuse (what a perfect implementation of this design streams):
    every depth map once (the gathers touch ~all of its sectors), every mask word once, the counts, the points once in and
    the prepared tiles once out + once in (k_cull_prepare -> k_frame_cull9; the 63 re-reads per 32-frame chunk come from L2),
    the poses.  ncu's dram__bytes (roofline.traffic) is compared with this number."""
    tiles = (n + TILE_POINTS - 1) // TILE_POINTS
    return (frames * DEPTH_BYTES + frames * stride_words * 4.0 + frames * n_objects * 4.0 + 12.0 * n + 2.0 * n
            + 2.0 * tiles * TILE_PREP_BYTES + frames * 128.0)


def workload_sizes(args):
    rng = np.random.default_rng(1000)
    lo, hi = int(args.points * 0.75), int(args.points * 1.25)
    return [int(v) for v in rng.integers(lo, hi + 1, args.scenes)]


def workload_config(args):
    """Identical in both arms (the driver compares it); run facts live in `run`, not here."""
    return {
        "workload": (f"BASELINE configs[1]: ScanNet visual_correspondence, {args.scenes} scenes x {args.pairs_per_scene} "
                     f"frame-pairs, ~{args.points // 1000}k pts/scene (N~U[0.75,1.25]x), independent-pair model "
                     f"(2 frames/pair: {2 * args.pairs_per_scene} poses + 640x480 u16 depth maps per scene), "
                     "points float32 xyz, synthetic room scenes"),
        "scenes": args.scenes, "pairs_per_scene": args.pairs_per_scene, "points_mean": args.points,
        "frames_per_scene": 2 * args.pairs_per_scene,
        "depth": "640x480 u16, rendered once at setup by this repo's z-buffer definition (+1 mm); parity of that render is unpinned (not in the reference)",
        "l2": "inputs larger than L2: every step streams all scenes' depth maps + masks (>> 126 MB), no flush needed",
    }


# ------------------------------------------------------------------------------------------------
# clocks sampler (nvidia-smi while the timed region runs)
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    """nvidia-smi -lms 100 in the background; samples are kept with their timestamps so that only those taken while the
    GPU runs the timed work are used (the tool needs ~1 s to start, the timed region of a default run lasts 0.25 s)."""
    FIELDS = ("timestamp,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu = gpu_index
        self.proc = None
        self.path = None

    def start(self):
        try:
            fd, self.path = tempfile.mkstemp(suffix=".csv")
            os.close(fd)
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.gpu}", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def samples_between(self, t0: float, t1: float) -> int:
        """Number of samples nvidia-smi has written with a timestamp inside [t0, t1] (wall-clock seconds)."""
        return len(self._read(t0, t1)[0])

    def _read(self, t0=None, t1=None):
        import datetime
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        try:
            for ln in open(self.path):
                p = [x.strip() for x in ln.split(",")]
                if len(p) < 8:
                    continue
                try:
                    ts = datetime.datetime.strptime(p[0], "%Y/%m/%d %H:%M:%S.%f").timestamp()
                    c, m = float(p[1]), float(p[2])
                except ValueError:
                    continue
                if t0 is not None and not (t0 <= ts <= t1):
                    continue
                sm.append(c); mx.append(m)
                for nm, v in zip(names, p[4:8]):
                    if v.lower().startswith("active"):
                        reasons.add(nm)
        except Exception:
            pass
        return sm, mx, reasons

    def stop(self, t0=None, t1=None, note=""):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.proc is None:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = self._read(t0, t1)
        try:
            os.unlink(self.path)
        except Exception:
            pass
        if sm:
            out.update(sm_mhz=float(np.median(sm)), sm_max_mhz=float(max(mx)), reasons=sorted(reasons), samples=len(sm))
        if note:
            out["note"] = note
        return out


def host_cores() -> int:
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


# ------------------------------------------------------------------------------------------------
# CPU arm.  kind "reference": the UNMODIFIED reference functions (oracle/ref_arm.py drives them from oracle/_ref or
# /root/reference); kind "port": the NumPy oracle port, only when no copy of the reference can be imported.
# ------------------------------------------------------------------------------------------------
def _cpu_pair_worker(job):
    """Port fallback: `n_pairs` independent pairs of one synthetic scene through oracle.geometry (data in RAM)."""
    seed, n_points, n_pairs = job
    try:
        from threadpoolctl import threadpool_limits
        threadpool_limits(1)
    except Exception:
        pass
    from multi_spatialmllm_b200 import synth
    from oracle import geometry as og
    s = synth.make_scene(seed=seed, n_points=n_points, n_frames=2 * n_pairs, n_objects=12)
    pts = s.points.astype(np.float64)
    H, W = s.image_hw
    rng = np.random.default_rng(seed)
    E = [og.aligned_pose(s.axis_align, s.c2w[f]) for f in range(s.n_frames)]
    depth = [og.depth_to_u16(og.zbuffer_render(pts, s.K, E[f], H, W, *s.depth_hw), 1) for f in range(s.n_frames)]
    t0 = time.perf_counter()
    done = 0
    for k in range(n_pairs):
        a, b = 2 * k, 2 * k + 1
        _, _, ma = og.frame_visibility(pts, s.K, E[a], depth[a], H, W)
        _, _, mb = og.frame_visibility(pts, s.K, E[b], depth[b], H, W)
        og.pair_relations(E[a], E[b], ma, mb)
        og.relative_displacement(E[a], E[b])
        common = og.common_visible_points(np.flatnonzero(ma), np.flatnonzero(mb))
        if len(common):
            p = int(common[int(rng.integers(0, len(common)))])
            og.annotate_point(pts[p], s.K, E[a], depth[a], H, W)
            og.annotate_point(pts[p], s.K, E[b], depth[b], H, W)
        og.object_visible_counts(ma, s.instance_ids, s.num_objects)
        og.object_visible_counts(mb, s.instance_ids, s.num_objects)
        done += 1
    return done, time.perf_counter() - t0


def _cpu_noop(_):
    import numpy  # noqa: F401
    from oracle import geometry  # noqa: F401
    return 0


def run_cpu_sample(n_points: int, pairs_per_worker: int, workers: int, seed0: int = 0, as_shipped: bool = False):
    """One bounded CPU sample -> dict(kind, pairs, busy, wall, [as_shipped_pairs, as_shipped_busy])."""
    from oracle import ref_shim
    if ref_shim.reference_available():
        from oracle import ref_arm
        variants = ("io_free", "as_shipped") if as_shipped else ("io_free",)
        res, wall = ref_arm.run_sample(n_points, pairs_per_worker, workers, seed0=seed0, variants=variants)
        out = {"kind": "reference", "pairs": res["io_free"][0], "busy": res["io_free"][1], "wall": wall}
        if as_shipped:
            out.update(as_shipped_pairs=res["as_shipped"][0], as_shipped_busy=res["as_shipped"][1])
        return out
    import multiprocessing as mp
    ctx = mp.get_context("spawn")
    jobs = [(seed0 + w, n_points, pairs_per_worker) for w in range(workers)]
    with ctx.Pool(workers) as pool:
        pool.map(_cpu_noop, range(workers))
        t0 = time.perf_counter()
        res = pool.map(_cpu_pair_worker, jobs)
        wall = time.perf_counter() - t0
    return {"kind": "port", "pairs": sum(r[0] for r in res), "busy": max(r[1] for r in res), "wall": wall}


def cpu_sample_text(kind, workers, ppw, points):
    code = ("the UNMODIFIED reference functions (SceneInfoHandler.project_3d_point_to_image + check_point_visibility per frame, "
            "extract_yaw_pitch, calculate_camera_overlap, np.intersect1d, 2x get_point_2d_coordinates_in_image, per-object set "
            "intersections), I/O-free variant: depth maps, image shape and points served from RAM by a handler subclass"
            if kind == "reference" else
            "NumPy oracle port of IH:46-72,337-386 / CFR:102-137 / VC-coor:303-344 / COV:133-136 (no copy of the reference importable)")
    return f"{workers} processes x {ppw} independent pairs of a {points}-point scene per step; {code}"


def main_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    cores = host_cores()
    workers = max(1, min(cores, 64))      # more workers than ~64 lowered the aggregate rate on the GPU host (memory-bound NumPy)
    ppw = max(1, args.cpu_pairs_per_worker // 3)
    times, pairs_tot, kind, shipped = [], 0, "port", None
    for it in range(args.warmup + args.steps):
        first_timed = it == args.warmup
        r = run_cpu_sample(args.points, ppw, workers, seed0=1000 * it, as_shipped=first_timed)
        kind = r["kind"]
        if it >= args.warmup:
            times.append(r["busy"])
            pairs_tot += r["pairs"]
        if first_timed and "as_shipped_busy" in r:
            shipped = r["as_shipped_pairs"] / r["as_shipped_busy"]
    total = float(sum(times))
    value = pairs_tot / total
    cb = {"value": value, "unit": UNIT, "cores": workers, "host_cores": cores, "kind": kind,
          "sample": cpu_sample_text(kind, workers, ppw, args.points)}
    if shipped is not None:
        cb["as_shipped_value"] = shipped
        cb["as_shipped_note"] = ("same pairs through the stock handler: every call decodes the JPEG (shape) and the PNG (depth) and every "
                                 "per-point call reloads aligned_points.npy (IH:133-139,153-157,272-276); first timed step only")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * total / max(args.steps, 1), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(args), "cpu_baseline": cb,
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "run": {"note": "CPU arm: every step is a bounded sample of the workload named in config"},
    }
    print(json.dumps(line), file=args.out, flush=True)
    return 0


# ------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------
def parity_check(sc, s, P):
    """Before the clock starts: two frames' masks and five pair records of scene 0 against the CPU oracle."""
    from multi_spatialmllm_b200.engine import DeviceScene
    from oracle import geometry as og
    pts = s.points.astype(np.float64)
    H, W = s.image_hw
    depth = sc.depth[:12].cpu().numpy()
    E = [og.aligned_pose(s.axis_align, s.c2w[f]) for f in range(10)]
    w = sc.mask[:10].cpu().numpy().view(np.uint32)
    bits = np.unpackbits(w.view(np.uint8), axis=1, bitorder="little")[:, :sc.n].astype(bool)
    if sc.inv_perm is not None:
        bits = bits[:, sc.inv_perm.cpu().numpy()]
    masks = {}
    for f in range(10):
        masks[f] = og.frame_visibility(pts, s.K, E[f], depth[f], H, W)[2]
    for f in (0, 1):
        if not np.array_equal(masks[f], bits[f]):
            raise AssertionError(f"parity check failed: visibility mask of frame {f} of scene 0 differs from the oracle")
    rec = DeviceScene.records_to_numpy(sc.rec[:5])
    cnt = sc.counts[:10].cpu().numpy()
    for k in range(5):
        a, b = 2 * k, 2 * k + 1
        r = rec[k]
        if not (np.array_equal(bits[a], masks[a]) and np.array_equal(bits[b], masks[b])):
            raise AssertionError(f"parity check failed: masks of pair {k}")
        common = np.flatnonzero(masks[a] & masks[b])
        if r["inter"] != len(common) or r["uni"] != int((masks[a] | masks[b]).sum()):
            raise AssertionError(f"parity check failed: overlap counts of pair {k}")
        if len(common) == 0:
            if r["point_id"] != -1:
                raise AssertionError(f"parity check failed: pair {k} has no common point")
            continue
        if not (r["point_id"] in common):
            raise AssertionError(f"parity check failed: sampled point of pair {k} is not commonly visible")
        for slot, f in ((0, a), (1, b)):
            an = og.annotate_point(pts[r["point_id"]], s.K, E[f], depth[f], H, W)
            if an is None or [int(r["permille"][2 * slot]), int(r["permille"][2 * slot + 1])] != [an["x_permille"], an["y_permille"]] \
                    or int(r["depth_mm"][slot]) != an["depth_mm"]:
                raise AssertionError(f"parity check failed: annotation of pair {k}, frame {f}")
        for f in (a, b):
            if not np.array_equal(cnt[f], og.object_visible_counts(masks[f], s.instance_ids, s.num_objects)):
                raise AssertionError(f"parity check failed: object counts of frame {f}")
    return True


def main_gpu(args):
    import ctypes as C

    import torch
    import torch.distributed as dist

    from multi_spatialmllm_b200 import _lib, synth  # raises if libmsp.so is missing: no fallback
    from multi_spatialmllm_b200.engine import DeviceScene, _ptr, _stream, check, lib, mask_stride_words
    from multi_spatialmllm_b200.pipeline import HostScene, PairAnnotator

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    assert torch.cuda.is_available(), "bench.py needs a CUDA device; use --impl reference for the CPU arm"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return float(x)
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(x):
        if world == 1:
            return float(x)
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    sizes = workload_sizes(args)
    P = args.pairs_per_scene
    F = 2 * P
    # ---- fit the resident set into HBM (depth dominates: F * 614400 B per scene)
    free_b, total_b = torch.cuda.mem_get_info()
    per_scene = F * DEPTH_BYTES + F * (sum(sizes) / len(sizes) / 8 + 64) + 16 * sum(sizes) / len(sizes) + F * 500
    n_scenes = args.scenes
    while n_scenes > 1 and n_scenes * per_scene > 0.86 * free_b:
        n_scenes -= 1
    sizes = sizes[:n_scenes]

    # ---- setup (untimed): scenes in HBM, depth maps rendered by K2, pair lists, selection words
    def build_scene(i, n, frames, pairs, reorder=None, shuffle=False, seed_base=0):
        s = synth.make_scene(seed=seed_base + rank * 100003 + i, n_points=n, n_frames=frames, n_objects=12, shuffle_points=shuffle)
        sc = DeviceScene(s.points, s.K, s.c2w, axis_align=s.axis_align, instance_ids=s.instance_ids,
                         num_objects=s.num_objects, image_hw=s.image_hw, depth_hw=s.depth_hw, device=dev, reorder=reorder)
        sc.render_depth(margin_mm=1)
        sc.c2w_in, sc.A_in = sc.c2w, torch.from_numpy(s.axis_align.reshape(16)).to(dev)
        sc.pairs = pairs if pairs is not None else torch.arange(frames, dtype=torch.int32, device=dev).view(-1, 2)
        npair = int(sc.pairs.shape[0])
        g = torch.Generator(device=dev); g.manual_seed(1234 + i)
        sc.sel = torch.randint(-2**31, 2**31 - 1, (npair,), dtype=torch.int32, device=dev, generator=g)
        sc.rec = torch.empty((npair, 128), dtype=torch.uint8, device=dev)
        sc.counts = torch.empty((frames, sc.num_objects), dtype=torch.int32, device=dev)
        sc.mask = torch.empty((frames, sc.stride), dtype=torch.int32, device=dev)   # every word is written by K1
        return s, sc

    scenes = []
    t_setup = time.time()
    s0 = None
    for i, n in enumerate(sizes):
        s, sc = build_scene(i, n, F, None, reorder=args.reorder)
        if i == 0:
            s0 = s
        scenes.append(sc)
    torch.cuda.synchronize()
    t_setup = time.time() - t_setup
    launches_per_scene = 4      # pose prologue, k_cull_prepare, k_frame_cull9 (+fused object counts), pair stage (+1 memset node)

    def scene_step(sc, ev=None, sel_mode=1):
        st = _stream()
        npair = int(sc.pairs.shape[0])
        check(lib.msp_pose_prepare(_ptr(sc.c2w_in), _ptr(sc.A_in), sc.F, _ptr(sc.E), _ptr(sc.w2c), _ptr(sc.yawpitch),
                                   _ptr(sc.valid), st))
        if ev is not None:
            ev[0].record()
        # masks + per-(frame, object) counts in one pass over the points (K1 with fused COV counting)
        check(lib.msp_project_visibility_objects(_ptr(sc.pts), sc.n, sc.pts_dtype, sc.K_host.ctypes.data, _ptr(sc.w2c),
                                                 sc.F, _ptr(sc.depth), C.byref(sc.cam), _ptr(sc.mask), sc.stride,
                                                 _ptr(sc.instance_ids), sc.num_objects, _ptr(sc.counts), st))
        if ev is not None:
            ev[1].record()
        check(lib.msp_pair_annotate(_ptr(sc.mask), sc.n, sc.stride, _ptr(sc.pairs), npair, _ptr(sc.sel), sel_mode, _ptr(sc.pts),
                                    sc.pts_dtype, sc.K_host.ctypes.data, _ptr(sc.E), _ptr(sc.w2c), _ptr(sc.yawpitch),
                                    _ptr(sc.depth), C.byref(sc.cam), _ptr(sc.rec), st))

    def step(ev=None):
        for k, sc in enumerate(scenes):
            scene_step(sc, None if ev is None else ev[k])

    sampler = ClockSampler(local)
    sampler.start()
    for _ in range(args.warmup):
        step()
    barrier()
    # ---- parity of the TIMED workload against the oracle before the clock starts (abort on any difference)
    parity_checked = parity_check(scenes[0], s0, P) if (rank == 0 and args.reorder is None) else None
    r0 = DeviceScene.records_to_numpy(scenes[0].rec)
    assert (r0["uni"] > 0).all() and (r0["point_id"] >= 0).sum() > 0.5 * P, "step produced empty records"

    evs = [[(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in scenes]
           for _ in range(args.steps)]
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    w0 = time.time()
    e0.record()
    for it in range(args.steps):
        step(evs[it])
    e1.record()
    barrier()
    # clocks DURING the timed work: nvidia-smi (started before the warm-up) samples every 100 ms; a default run's timed region is
    # shorter than that, so the same steps keep running, untimed, until at least three samples fall inside the loaded window
    extra_steps = 0
    while sampler.samples_between(w0, time.time()) < 3 and time.time() - w0 < 6.0:
        step()
        torch.cuda.synchronize()
        extra_steps += 1
    w1 = time.time()
    clocks = sampler.stop(w0, w1, note=f"sampled while the timed steps ran and {extra_steps} identical untimed steps right after them")
    elapsed_local = e0.elapsed_time(e1)
    k1_ms = sum(a.elapsed_time(b) for step_ev in evs for (a, b) in step_ev)
    elapsed_ms = max_over_ranks(elapsed_local)
    total_scenes = int(sum_over_ranks(len(scenes)))   # host-side manifest gather stand-in: total scenes done
    pairs_total = total_scenes * P * args.steps
    value = pairs_total / (elapsed_ms * 1e-3)

    # ---- roofline of the dominant kernel (K1 = k_cull_prepare + k_frame_cull9 behind msp_project_visibility_objects), rank 0
    k1_launches = len(scenes) * args.steps
    comp_bytes = sum(compulsory_bytes_launch(sc.n, sc.F, sc.stride, sc.num_objects) for sc in scenes) * args.steps
    alg_bytes = sum(sc.F * algorithmic_bytes_frame(sc.n) for sc in scenes) * args.steps
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    achieved = comp_bytes / (k1_ms * 1e-3) / 1e9
    traffic = traffic_note = None
    try:
        _tr = json.load(open(os.path.join(ROOT, "profiles", "k1_traffic.json")))
        traffic, traffic_note = _tr.get("dram_bytes_per_launch"), _tr.get("note")
    except Exception:
        pass
    roofline = {"bound": "hbm", "kernel": "k_frame_cull9 (+ k_cull_prepare) = msp_project_visibility_objects", "achieved": achieved,
                "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "peak_source": "MEASURED_PEAKS.json hbm_gbs (of measured)" if peaks else "fallback 6650 GB/s (of fallback)",
                "traffic": traffic, "traffic_note": traffic_note, "compulsory_bytes_per_launch": comp_bytes / k1_launches,
                "avg_launch_ms": k1_ms / k1_launches, "share_of_step": k1_ms / elapsed_local,
                "bytes_model": ("compulsory HBM bytes WITH the kernel's reuse: depth maps + mask words + counts + points in + prepared "
                                "tiles out/in + poses, each once per launch (bench.py compulsory_bytes_launch)"),
                "algorithmic_bytes_per_launch": alg_bytes / k1_launches,
                "algorithmic_frac": alg_bytes / (k1_ms * 1e-3) / 1e9 / peak,
                "algorithmic_note": ("SURVEY.md §8(d)'s independent-pair bytes (every frame streams the scene's points: 12N + 614400 + 80 + "
                                     "N/8).  NOT a physical fraction: the kernel stages each 1024-point tile once per 32-frame chunk and "
                                     "re-reads come from L2, so this ratio can exceed 1; `frac` is the physical one")}
    if traffic:
        dram_gbs = traffic / (k1_ms / k1_launches * 1e-3) / 1e9
        roofline.update(dram_gbs=dram_gbs, dram_frac=dram_gbs / peak,
                        traffic_note="ncu dram__bytes_read+write of one launch on a 200k-point x 2000-frame scene (profiles/k1_traffic.json)"
                                     + ("; " + traffic_note if traffic_note else ""))

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": elapsed_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32/f64", "data": "synthetic", "config": workload_config(args), "roofline": roofline,
        "gpu_launches": launches_per_scene * len(scenes) * args.steps, "clocks": clocks,
        "parity_checked": parity_checked,
        "run": {"scenes_resident_per_gpu": len(scenes), "setup_s": round(t_setup, 1), "reorder": args.reorder,
                "parity": ("scene 0 of the timed workload vs oracle.geometry before the clock: masks of frames 0-9, 5 pair records "
                           "(overlap counts, sampled point, per-mille / mm annotation), object counts of 10 frames"
                           if parity_checked else "skipped (reordered scene or rank > 0)")},
    }

    # ---- e2e through the public API with HOST buffers (rank-local; same metric)
    if not args.no_e2e:
        ne = min(args.e2e_scenes, len(scenes))
        hosts = []
        for i in range(ne):
            sc = scenes[i]
            s = synth.make_scene(seed=rank * 100003 + i, n_points=sc.n, n_frames=F, n_objects=12)
            hs = HostScene(torch.from_numpy(s.points), s.K, s.axis_align, torch.from_numpy(s.c2w),
                           sc.depth.cpu(), torch.from_numpy(s.instance_ids), s.num_objects, s.image_hw, s.depth_hw).pin()
            hs.pairs = sc.pairs.cpu().pin_memory()
            hs.sel = sc.sel.cpu().pin_memory()
            hosts.append(hs)
        ann = PairAnnotator(dev, n_streams=2)

        def e2e_step():
            hb = db = 0
            for hs in hosts:
                h = ann.submit(hs, hs.pairs, hs.sel, sel_mode=1)
                hb += h["h2d_bytes"]; db += h["d2h_bytes"]
            out = ann.drain()
            return hb, db, out

        for _ in range(max(1, min(args.warmup, 2))):
            hb, db, out = e2e_step()
        # the public API returns the same records the resident path computed
        ref = DeviceScene.records_to_numpy(scenes[0].rec)
        assert np.array_equal(out[0]["records"]["point_id"], ref["point_id"]) and np.array_equal(out[0]["records"]["inter"], ref["inter"])
        # plain pinned -> device copy rate on this rank WHILE every rank copies: the PCIe / host-memory ceiling of the e2e number
        pin = hosts[0].depth
        dst = torch.empty_like(pin, device=dev)
        barrier()
        c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        c0.record()
        for _ in range(3):
            dst.copy_(pin, non_blocking=True)
        c1.record()
        torch.cuda.synchronize()
        h2d_peak = 3 * pin.numel() * 2 / (c0.elapsed_time(c1) * 1e-3) / 1e9
        del dst
        barrier()
        t0 = time.perf_counter()
        esteps = max(1, min(args.steps, args.e2e_steps))
        for _ in range(esteps):
            hb, db, out = e2e_step()
        torch.cuda.synchronize()
        dt_local = time.perf_counter() - t0
        dt = max_over_ranks(dt_local)
        line["e2e"] = {"value": world * ne * P * esteps / dt, "unit": UNIT, "h2d_bytes_per_step": hb, "d2h_bytes_per_step": db,
                       "scenes_per_step": ne, "steps": esteps,
                       "h2d_gbs_this_rank": hb * esteps / dt_local / 1e9, "h2d_memcpy_peak_gbs_this_rank": h2d_peak,
                       "limiter": ("PCIe: every pair ships two raw 640x480 u16 depth maps (1.23 MB); h2d_gbs_this_rank vs the plain "
                                   "cudaMemcpyAsync rate measured while all ranks copy (h2d_memcpy_peak_gbs_this_rank)"),
                       "api": "multi_spatialmllm_b200.pipeline.PairAnnotator.submit/drain on pinned host buffers"}
        del hosts, ann

    # ---- the other BASELINE configs, small samples.  The main workload's scenes are released first.
    if not args.no_extra:
        del scenes
        torch.cuda.empty_cache()
        line["extra"] = extra_configs(args, dev, rank, world, build_scene, scene_step, barrier, max_over_ranks, sum_over_ranks)

    # ---- CPU baseline on this box's host cores (rank 0, N=1 only; bounded sample)
    if rank == 0 and world == 1 and not args.no_cpu:
        cores = host_cores()
        workers = max(1, min(cores, 64))
        r = run_cpu_sample(args.points, args.cpu_pairs_per_worker, workers)
        line["cpu_baseline"] = {
            "value": r["pairs"] / r["busy"], "unit": UNIT, "cores": workers, "host_cores": cores, "kind": r["kind"],
            "sample": cpu_sample_text(r["kind"], workers, args.cpu_pairs_per_worker, args.points)
            + f" ({r['pairs']} pairs, {r['busy']:.1f} s busy, {r['wall']:.1f} s wall incl. scene setup)"}

    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if rank == 0:
        print(json.dumps(line), file=args.out, flush=True)
    return 0


def extra_configs(args, dev, rank, world, build_scene, scene_step, barrier, max_over_ranks, sum_over_ranks):
    """Small timed samples of the configurations the headline does not cover; every number is device-timed (CUDA events,
    max over ranks) on resident inputs unless it says e2e."""
    import ctypes as C

    import torch

    from multi_spatialmllm_b200 import synth
    from multi_spatialmllm_b200.engine import DeviceScene, _ptr, _stream, check, lib
    from multi_spatialmllm_b200.pipeline import HostScene, PairAnnotator
    out = {}
    steps = 3

    def timed(fn, reps):
        fn()                                                  # warm-up
        barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(reps):
            fn()
        b.record()
        barrier()
        return max_over_ranks(a.elapsed_time(b)) / reps

    # ---- (i) BASELINE configs[4] "synthetic stress": 500k uniform random points, 640x480 depth RENDERED INSIDE the step (K2 +
    #          u16 conversion) + correspondence (K1 masks + pair stage), independent pairs; point order as given and Morton-sorted
    n_st, f_st, sc_st = 500_000, 128, 2
    for reorder in (None, "morton"):
        scs, sort_ms = [], 0.0
        for i in range(sc_st):
            s = synth.make_stress_scene(seed=7000 + rank * 131 + i, n_points=n_st, n_frames=f_st)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            sc = DeviceScene(s.points, s.K, s.c2w, axis_align=s.axis_align, image_hw=s.image_hw, depth_hw=s.depth_hw,
                             device=dev, reorder=reorder)
            torch.cuda.synchronize()
            sort_ms += (time.perf_counter() - t0) * 1e3
            sc.pairs = torch.arange(f_st, dtype=torch.int32, device=dev).view(-1, 2)
            sc.sel = torch.randint(-2**31, 2**31 - 1, (f_st // 2,), dtype=torch.int32, device=dev)
            scs.append(sc)
        ev = {k: [torch.cuda.Event(enable_timing=True) for _ in range(4)] for k in range(sc_st)}

        def stress_step():
            for k, sc in enumerate(scs):
                ev[k][0].record()
                sc.render_depth(margin_mm=1, chunk=f_st, keep_buffers=True)   # K2: z-buffer splat + float32 -> u16 mm
                ev[k][1].record()
                sc.project_visibility()                            # K1 (no instances in this config)
                ev[k][2].record()
                sc.pair_annotate(sc.pairs, sc.sel, sel_mode=1)     # correspondence: overlap + sampled common point in both frames
                ev[k][3].record()
        ms = timed(stress_step, steps)
        torch.cuda.synchronize()
        k2 = sum(ev[k][0].elapsed_time(ev[k][1]) for k in ev) / sc_st
        k1 = sum(ev[k][1].elapsed_time(ev[k][2]) for k in ev) / sc_st
        pairs = sum_over_ranks(sc_st * f_st // 2)
        key = "stress_500k_morton" if reorder else "stress_500k_as_given"
        out[key] = {"pairs_per_s": pairs / (ms * 1e-3), "ms_per_step": ms, "k2_render_ms_per_scene": k2, "k1_ms_per_scene": k1,
                    "frames_per_scene": f_st, "scenes_per_gpu": sc_st, "points": n_st,
                    "upload_and_sort_ms_per_scene": sort_ms / sc_st,
                    "note": ("depth render inside the timed step; upload_and_sort is the one-time DeviceScene construction (host->device "
                             "copy" + (" + 30-bit Morton sort" if reorder else "") + "), not in ms_per_step")}
        del scs
        torch.cuda.empty_cache()

    # ---- (ii) configs[1] with the point order destroyed (shuffle_points=True), as given and Morton-sorted
    for reorder in (None, "morton"):
        scs = [build_scene(i, args.points, 2 * args.pairs_per_scene, None, reorder=reorder, shuffle=True, seed_base=555)[1] for i in range(2)]
        ev2 = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in scs]

        def shuf_step():
            for k, sc in enumerate(scs):
                scene_step(sc, ev2[k])
        ms = timed(shuf_step, steps)
        torch.cuda.synchronize()
        k1 = sum(a.elapsed_time(b) for a, b in ev2) / len(scs)
        pairs = sum_over_ranks(len(scs) * args.pairs_per_scene)
        out["configs1_shuffled_morton" if reorder else "configs1_shuffled_as_given"] = {
            "pairs_per_s": pairs / (ms * 1e-3), "k1_ms_per_launch": k1, "scenes_per_gpu": len(scs),
            "note": "same scenes as the headline but with a random point permutation; fused object counts on (every 64-point group mixes ids)"}
        del scs
        torch.cuda.empty_cache()

    # ---- (ii-b) configs[1] with a ScanNet-like instance count (counting fused for any number of objects)
    def build_many_objects(i, n_obj):
        s = synth.make_scene(seed=4242 + rank * 100003 + i, n_points=args.points, n_frames=2 * args.pairs_per_scene, n_objects=n_obj)
        sc = DeviceScene(s.points, s.K, s.c2w, axis_align=s.axis_align, instance_ids=s.instance_ids, num_objects=s.num_objects,
                         image_hw=s.image_hw, depth_hw=s.depth_hw, device=dev)
        sc.render_depth(margin_mm=1)
        sc.c2w_in, sc.A_in = sc.c2w, torch.from_numpy(s.axis_align.reshape(16)).to(dev)
        sc.pairs = torch.arange(sc.F, dtype=torch.int32, device=dev).view(-1, 2)
        sc.sel = torch.randint(-2**31, 2**31 - 1, (sc.F // 2,), dtype=torch.int32, device=dev)
        sc.rec = torch.empty((sc.F // 2, 128), dtype=torch.uint8, device=dev)
        sc.counts = torch.empty((sc.F, sc.num_objects), dtype=torch.int32, device=dev)
        sc.mask = torch.empty((sc.F, sc.stride), dtype=torch.int32, device=dev)
        return sc
    for n_obj in (12, 150):
        scs = [build_many_objects(i, n_obj) for i in range(2)]
        ev3 = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in scs]

        def obj_step():
            for k, sc in enumerate(scs):
                scene_step(sc, ev3[k])
        ms = timed(obj_step, steps)
        torch.cuda.synchronize()
        out[f"configs1_{n_obj}_objects"] = {"k1_ms_per_launch": sum(a.elapsed_time(b) for a, b in ev3) / len(scs),
                                           "pairs_per_s": sum_over_ranks(len(scs) * args.pairs_per_scene) / (ms * 1e-3),
                                           "note": "one launch: masks + per-(frame, object) counts, any object count (per-tile id table)"}
        del scs
        torch.cuda.empty_cache()

    # ---- (iii) configs[3]-shaped shared-frame scene: N = 200k, F = 375, ALL 70 125 pairs (CFR:176-189), resident and e2e
    f_sh = 375
    iu = np.triu_indices(f_sh, 1)
    pairs_np = np.stack(iu, 1).astype(np.int32)
    pairs_sh = torch.from_numpy(pairs_np).to(dev)
    s_sh, sc_sh = build_scene(0, args.points, f_sh, pairs_sh, seed_base=999)
    ms = timed(lambda: scene_step(sc_sh, None, sel_mode=1), steps)
    npairs = int(pairs_sh.shape[0])
    out["shared_frame_375"] = {"pairs_per_s": sum_over_ranks(npairs) / (ms * 1e-3), "ms_per_scene": ms, "frames": f_sh, "pairs": npairs,
                               "points": args.points,
                               "note": "one scene, all i<j pairs of its 375 frames: pose prologue + K1 + pair stage (overlap, relations, common point)"}
    hs = HostScene(torch.from_numpy(s_sh.points), s_sh.K, s_sh.axis_align, torch.from_numpy(s_sh.c2w), sc_sh.depth.cpu(),
                   torch.from_numpy(s_sh.instance_ids), s_sh.num_objects, s_sh.image_hw, s_sh.depth_hw).pin()
    hp, hsel = pairs_sh.cpu().pin_memory(), sc_sh.sel.cpu().pin_memory()
    ann = PairAnnotator(dev, n_streams=2)
    for _ in range(2):
        ann.submit(hs, hp, hsel, sel_mode=1)
    ann.drain()
    barrier()
    t0 = time.perf_counter()
    for _ in range(4):
        h = ann.submit(hs, hp, hsel, sel_mode=1)
    ann.drain()
    dt = max_over_ranks(time.perf_counter() - t0)
    out["shared_frame_375"]["e2e_pairs_per_s"] = sum_over_ranks(4 * npairs) / dt
    out["shared_frame_375"]["e2e_h2d_bytes_per_scene"] = h["h2d_bytes"]
    out["shared_frame_375"]["e2e_d2h_bytes_per_scene"] = h["d2h_bytes"]
    del sc_sh, ann, hs
    torch.cuda.empty_cache()

    # ---- (iv) the regeneration driver on configs[3]-shaped scenes: K1 + all-pairs relations + the exact
    #          sample_dataframe replica + the batch writers of four streams, JSONL written to a scratch directory (wall clock)
    import shutil
    from multi_spatialmllm_b200 import regen, streams
    tmp = tempfile.mkdtemp(prefix="msp_regen_")
    try:
        rng = np.random.default_rng(4321 + rank)
        n_sc, frames_done, pairs_done, samples = 2, 0, 0, 0
        scs = []
        for i in range(n_sc):
            s = synth.make_scene(seed=8800 + rank * 17 + i, n_points=args.points, n_frames=f_sh, n_objects=12)
            sc = DeviceScene(s.points, s.K, s.c2w, axis_align=s.axis_align, instance_ids=s.instance_ids, num_objects=s.num_objects, device=dev)
            sc.render_depth(1)
            scs.append((s, sc))
        regen.regenerate_scene(scs[0][1], scs[0][0].scene_id, scs[0][0].image_ids, 200, 32, np.random.default_rng(1))   # untimed: lazy imports
        barrier()
        t0 = time.perf_counter()
        for s, sc in scs:
            res, tabs = regen.regenerate_scene(sc, s.scene_id, s.image_ids, 200, 32, rng)
            frames_done += sc.F
            pairs_done += 0 if not tabs["relations"] else len(tabs["relations"]["pairs"])
            for k, v in res.items():
                samples += streams.write_jsonl(v, os.path.join(tmp, f"{k}.part-{rank:05d}.jsonl"))
        torch.cuda.synchronize()
        dt = max_over_ranks(time.perf_counter() - t0)
        out["regen"] = {"scenes_per_gpu": n_sc, "frames_per_scene": f_sh, "points": args.points,
                        "pairs_related_per_s": sum_over_ranks(pairs_done) / dt, "samples_per_s": sum_over_ranks(samples) / dt,
                        "frames_per_s": sum_over_ranks(frames_done) / dt, "seconds": dt,
                        "note": ("regen.regenerate_scene per scene: masks + counts, all-pairs relations, 200 pairs drawn by the reference's "
                                 "overlap-bin sampler, visual_correspondence + 9 camera_movement question types + depth_estimation / "
                                 "depth_comparison samples, JSONL to local scratch; wall clock, host formatting included")}
        del scs
    finally:
        shutil.rmtree(tmp, ignore_errors=True)
    torch.cuda.empty_cache()
    return out


def _claim_stdout():
    """Keep stdout for the ONE JSON line: libraries that print to fd 1 (NCCL's version banner, torchrun notices) are
    sent to stderr instead.  Returns a file object bound to the real stdout."""
    sys.stdout.flush()
    real = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    return real


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--scenes", type=int, default=100)
    ap.add_argument("--pairs-per-scene", type=int, default=1000)
    ap.add_argument("--points", type=int, default=200_000)
    ap.add_argument("--e2e-scenes", type=int, default=3)
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--cpu-pairs-per-worker", type=int, default=24)
    ap.add_argument("--reorder", default=None, choices=[None, "morton"], help="store scene points in Morton order (tuning)")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-extra", action="store_true")
    args = ap.parse_args()
    args.out = _claim_stdout()
    if args.impl == "reference":
        return main_reference(args)
    return main_gpu(args)


if __name__ == "__main__":
    sys.exit(main())
