#!/usr/bin/env python
"""Benchmark of the rasterization hot path (BASELINE.json metric) on 1..8 B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--config c1|c2|c3|c4|c5] [--phong] [--textured]
                    [--impl ours|reference]

A "step" is one frame: the whole hot path (setup -> bin -> raster) over one synthetic scene.
Default workload = BASELINE.json configs[1] ("C2": 1 M ~10-pixel triangles, 1920x1080,
depth-tested, Gouraud).  With N > 1 ranks (torchrun, one process per GPU) the path shards by
FRAME: every rank renders its own 1 M-triangle frame (weak scaling, no data-path collective);
the NCCL gather of the finished colour images to rank 0 is timed separately ("with_gather").
The other configurations: c1 = the reference's demo sphere (with a whole-object leg), c3 = 50 k
large triangles at 4K, c4 = 20 M triangles at 16K^2 split into row bands over the ranks (strong),
c5 = 256 views of a 2 M-triangle mesh split over the ranks (strong).  --phong / --textured select
the per-pixel Phong and the textured, perspective-correct path.  At least 3 warm-up steps always run.

Prints ONE JSON line (see the contract in the task statement): value = device-resident
throughput (CUDA events, max over ranks), e2e = same metric through the host-pointer C-ABI call
(H2D + kernels + D2H inside the timed region), roofline = dominant kernel vs measured HBM peak,
cpu_baseline = the verbatim reference scalar path on this box's host cores.

--impl reference times the reference's own CPU implementation (oracle/_ref, the verbatim
FillEdgeTable + DrawModel per single-triangle object, all host threads) on the same workload.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRICS = {
    "c1": ("Mpixels/s", "C1: the reference's demo, ConstructSphere (2 208 triangles) as one object, 1920x1080"),
    "c2": ("Mtriangles/s", "C2: 1M ~10px triangles, 1920x1080, depth-tested Gouraud (setup/binning bound)"),
    "c3": ("Mpixels/s", "C3: 50k large overlapping triangles, 3840x2160, ~35x overdraw (fill bound)"),
    "c4": ("Mpixels/s", "C4: 20M triangles, 16384x16384, screen-space tile bands across the GPUs, NCCL gather"),
    "c5": ("Mtriangles/s", "C5: 256 views of a 2M-triangle mesh (ConstructSphere, StepCount 708), 1920x1080, frame-parallel"),
}
# c2/c3: every rank renders its own frame (weak).  c4: one frame split in row bands, c5: a fixed
# set of 256 views split over the ranks (strong: the total work does not grow with N).
SCALING = {"c1": "weak", "c2": "weak", "c3": "weak", "c4": "strong", "c5": "strong"}
C5_VIEWS = 256             # cpu_renderer_b200.scene.C5_VIEWS
TEX_SIZE = 1024                # --textured: the objects' Bitmap is TEX_SIZE x TEX_SIZE ARGB8


def log(*a):
    print(*a, file=sys.stderr, flush=True)


# stdout carries exactly ONE line, the JSON result.  Libraries write to fd 1 behind Python's back
# (NCCL prints its version banner there when a process group comes up), so fd 1 is pointed at stderr
# for the whole run and the result goes to the saved descriptor.
_REAL_STDOUT = None


def claim_stdout():
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.dup(1)
        os.dup2(2, 1)


def emit_line(line):
    data = (json.dumps(line) + "\n").encode()
    if _REAL_STDOUT is None:
        sys.stdout.write(data.decode()); sys.stdout.flush()
    else:
        sys.stdout.flush()
        os.write(_REAL_STDOUT, data)


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def ncu_traffic(config, kernel):
    """DRAM bytes per launch of `kernel` from the committed ncu --set full capture, if any."""
    p = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(p):
        try:
            return json.load(open(p)).get(config, {}).get(kernel)
        except Exception:
            return None
    return None


class ClockSampler:
    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.tmp = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.QUERY}", "--format=csv,noheader,nounits",
                 "-lms", "100", "-i", str(gpu_index)], stdout=self.tmp, stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None
        self.marks = []

    def mark(self):
        self.marks.append(time.time())

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.proc is None:
            return out
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        self.tmp.flush()
        rows = [l.strip().split(", ") for l in open(self.tmp.name) if l.strip()]
        os.unlink(self.tmp.name)
        sm, reasons, mx = [], set(), None
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in rows:
            try:
                sm.append(float(r[1])); mx = float(r[2])
                for n, v in zip(names, r[4:8]):
                    if v.strip().lower() == "active":
                        reasons.add(n)
            except Exception:
                continue
        if sm:
            # samples under load = upper half (the sampler also sees idle gaps between passes)
            s = sorted(sm)
            out.update(sm_mhz=float(np.median(s[len(s) // 2:])), sm_max_mhz=mx,
                       reasons=sorted(reasons), samples=len(sm))
        return out


def build_scene(config, rank, scale=1.0):
    from cpu_renderer_b200 import scene as sc
    if config == "c5":
        step = max(4, int(round(708 * scale ** 0.5)))
        pos, col, nrm, uvs = sc.construct_sphere(step)
        return sc.sphere_scene(pos, col, nrm, uvs, 1920, 1080, 500.0, name="c5")
    if config == "c1":
        # the verbatim ConstructSphere mesh (projekt.cpp:4123), committed with the golden vectors
        m = np.load(os.path.join(ROOT, "tests", "golden", "sphere_mesh.npz"))
        return sc.sphere_scene(m["pos"], m["col"], m["nrm"], m["uvs"], 1920, 1080, 500.0, name="c1")
    cfg = dict(sc.CONFIGS[config])
    cfg["count"] = max(1, int(round(cfg["count"] * scale)))
    if config in ("c2", "c3"):
        cfg["seed"] = cfg["seed"] + 0x1000 * rank   # every rank renders its own frame
    return sc.triangle_soup(config, **cfg)


def c5_view(i):
    """View i of config C5 (cpu_renderer_b200.scene.c5_view)."""
    from cpu_renderer_b200 import scene as sc
    return sc.c5_view(i)


# ------------------------------------------------------------------------------ reference arm
def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_lib as ol
    from cpu_renderer_b200 import scene as sc
    unit, workload = METRICS[args.config]
    scene = build_scene(args.config, 0, args.scale)
    if args.textured:
        scene = sc.textured(scene, TEX_SIZE, TEX_SIZE, lo=0.05, hi=0.95)
    s, sample, threads = cpu_sample(args.config, scene, sc, args.ref_sample, args.ref_threads)
    kind = "reference" if ol.ref_available() else "port"
    res = time_cpu(ol, s, threads, args.steps, args.warmup, kind, args.phong)
    ms = res["ms_per_step"]
    units = sample if unit == "Mtriangles/s" else scene.width * scene.height * (sample / scene.triangle_count)
    value = units / (ms * 1e-3) / 1e6
    line = {
        "impl": "reference", "metric": unit, "value": value, "unit": unit, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
        "scaling": SCALING[args.config], "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload + (" -- per-pixel Phong shading" if args.phong else "") +
                               (f" -- textured, {TEX_SIZE}x{TEX_SIZE} ARGB8, perspective correct" if args.textured else ""),
                   "shading": ("phong" if args.phong else "gouraud") + ("+texture" if args.textured else ""),
                   "triangles": scene.triangle_count, "width": scene.width, "height": scene.height,
                   "triangles_per_step": sample, "whole_frame": sample == scene.triangle_count, "scale": args.scale},
        "cpu_baseline": {"value": value, "unit": unit, "cores": threads, "kind": kind,
                         "sample": res["sample"]},
        "e2e": {"value": value, "unit": unit, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "host": {"cpu_count": os.cpu_count(), "model": cpu_model()},
    }
    emit_line(line)


def cpu_sample(config, scene, sc, sample=0, threads=0):
    """Bounded CPU sample of a config: a prefix of the frame's triangle list (for c5: of view 0)."""
    import copy
    # c1/c2/c3: the whole frame (same configuration as the GPU arm); c4/c5: a stated prefix / stride sample
    default = {"c1": 1 << 30, "c2": 1 << 30, "c3": 1 << 30, "c4": 250_000, "c5": 250_000}[config]
    sample = min(scene.triangle_count, sample or default)
    if config == "c5":
        # a mesh is ordered: take every k-th triangle so the sample covers the whole sphere
        k = max(1, scene.triangle_count // sample)
        pick = (np.arange(sample) * k)[:, None] * 3 + np.arange(3)[None, :]
        pick = pick.reshape(-1)
        arrays = [np.ascontiguousarray(a[pick]) for a in (scene.positions, scene.colors, scene.normals, scene.uvs)]
    else:
        arrays = [a[:sample * 3] for a in (scene.positions, scene.colors, scene.normals, scene.uvs)]
    s = sc.Scene(scene.name, scene.width, scene.height, copy.copy(scene.transform), *arrays,
                 scene.object_p, scene.ambient, scene.lights, texture=scene.texture)
    if config == "c5":
        s.object_p, s.transform.distance_above_target = c5_view(0)
    # every worker owns a private colour/depth pair: 2 GiB each at 16384^2, so c4 uses few workers
    cap = 4 if config == "c4" else 1 << 30
    threads = threads or min(os.cpu_count() or 1, cap)
    return s, sample, threads


def cpu_model():
    try:
        for l in open("/proc/cpuinfo"):
            if l.startswith("model name"):
                return l.split(":", 1)[1].strip()
    except Exception:
        pass
    return "unknown"


def equivalent_zbuffer(ol, scene, tri_bytes, frame_ms, peak_gbs):
    """B = tri_bytes x triangles + 4 B x depth-tested fragments + 8 B x depth passes -- the
    z-buffer traffic of the REFERENCE's own algorithm for this frame (read the depth per fragment, write
    depth + colour per pass, projekt.cpp:525-529) -- divided by OUR frame time.  Not a roofline: the tiles
    live in shared memory and almost none of these bytes reach HBM; it is the figure the config label
    "fill / z-buffer bandwidth bound" refers to.  Fragments and passes are counted by the oracle in
    submission order on one thread; they do not depend on the shading mode."""
    import dataclasses
    plain = dataclasses.replace(scene, texture=None)
    st = ol.oracle_render(plain)["stats"]
    b = tri_bytes * scene.triangle_count + 4.0 * st["Fragments"] + 8.0 * st["DepthPasses"]
    gbs = b / (frame_ms * 1e-3) / 1e9
    return {"bytes": b, "fragments": st["Fragments"], "depth_passes": st["DepthPasses"], "GB/s": gbs,
            "of_measured_hbm_peak": gbs / peak_gbs,
            "what": "z-buffer traffic of the reference's own algorithm for this frame / our frame time (not HBM traffic)"}


def compare_with_oracle(ol, scene, color, depth, phong):
    """The rendered frame against the CPU oracle's of the same scene: pixels whose depth bits differ, pixels whose
    colour differs in any channel, the largest channel difference in LSB, and how many pixels took a clamped texel
    (where the reference reads outside its bitmap; golden colour is defined by the clamp there)."""
    want = ol.oracle_render(scene, phong=phong)
    zdiff = int((want["z"].view(np.uint32) != depth.view(np.uint32)).sum())
    ch = np.abs(want["color"].view(np.uint8).astype(np.int16) - color.view(np.uint8).astype(np.int16))
    per_pixel = ch.reshape(color.shape[0], color.shape[1], 4).max(axis=2)
    return {"pixels": int(color.size), "depth_pixels_differing": zdiff,
            "colour_pixels_differing": int((per_pixel != 0).sum()), "colour_max_lsb": int(per_pixel.max()),
            "texel_clamps": int(want["stats"].get("TexelClamps", 0)),
            "what": "end-to-end frame vs the CPU oracle (oracle/raster_oracle.c) on the same scene"}


def issue_roofline(config, fragments, raster_ms, sm_mhz):
    """The raster kernel's second roofline: instruction issue.  With tiles resident in shared memory the
    kernel moves few HBM bytes per fragment but executes instructions for every one of them, so the bound
    that matters is (thread-instructions executed) / (148 SMs x 4 schedulers x 32 lanes x SM clock).  The
    instruction counts come from the committed ncu capture of the same kernel on the same config
    (profiles/traffic.json: smsp__inst_executed.sum and the active threads per instruction); the duration is
    this run's own event-timed one."""
    p = os.path.join(ROOT, "profiles", "traffic.json")
    try:
        inst = json.load(open(p)).get(config, {}).get("raster_kernel_inst")
    except Exception:
        inst = None
    if not inst or not fragments:
        return None
    mhz = float(sm_mhz or 1965.0)
    thread_inst = inst["warp_instructions"] * inst["threads_per_instruction"]
    lane_slots = 148 * 4 * 32 * mhz * 1e6 * raster_ms * 1e-3
    return {"kernel": "raster_kernel", "warp_instructions": inst["warp_instructions"],
            "warp_instructions_per_fragment": inst["warp_instructions"] / fragments,
            "thread_instructions_per_fragment": thread_inst / fragments,
            "issue_slots_used": inst["warp_instructions"] / (148 * 4 * mhz * 1e6 * raster_ms * 1e-3),
            "lane_slots_used": thread_inst / lane_slots, "sm_mhz": mhz,
            "what": "ncu instruction counts of the committed capture over this run's raster_kernel time; fragments = "
                    "depth-tested fragments of the reference's algorithm (oracle count)"}


def avx_scenes(sc):
    """The two inputs the reference's multithreaded AVX path is timed on (it only
    handles textured + Phong convex objects that do not overlap on screen): the demo sphere at 1080p (config C1)
    and a 4K frame tiled with 10 x 5 such spheres (the 'C3-like scene of non-overlapping convex objects')."""
    from dataclasses import replace
    m = np.load(os.path.join(ROOT, "tests", "golden", "sphere_mesh.npz"))
    uv = np.clip(m["uvs"], 0.0, 1.0).astype(np.float32)       # that path fetches texels without a range check
    tex = sc.make_texture(256, 256)
    c1 = replace(sc.sphere_scene(m["pos"], m["col"], m["nrm"], uv, 1920, 1080, 500.0, name="c1_tex_phong"), texture=tex)
    grid = replace(sc.sphere_scene(m["pos"], m["col"], m["nrm"], uv, 3840, 2160, 500.0, name="convex_grid_4k"), texture=tex)
    step = 360.0 / (500.0 * 2.0 / 3.0)                          # 360 px between centres, spheres are ~334 px wide
    ps = [((i - 4.5) * step, (j - 2.0) * step, 0.0) for j in range(5) for i in range(10)]
    return [("c1_textured_phong", c1, [(0.0, 0.0, 0.0)]), ("convex_grid_4k_textured_phong", grid, ps)]


def avx_baseline(renderer=None, threads=0, reps=5):
    """cpu_baseline.avx_mt: DrawModelOptimizedLines + FillLinesOptimized (projekt.cpp:3362-3613, 629-1490) with
    `threads` workers behind Platform.AddEntry, best and median of `reps` frames; beside it the same frames
    through b200r_render_objects (host buffers in and out) when a renderer is given."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_lib as ol
    from cpu_renderer_b200 import scene as sc
    if not ol.avx_available():
        return {"unavailable": "oracle/_ref/libprojekt_avx.so has not been built"}
    threads = threads or (os.cpu_count() or 1)
    out = {"kind": "reference", "path": "FillEdgeTable + DrawModelOptimizedLines on the submitting thread, FillLinesOptimized "
                                        "on the workers (oracle/_ref/libprojekt_avx.so: the reference text through 5 sed rules)",
           "cores": threads, "host": cpu_model(), "unit": "Mpixels/s", "scenes": {}}
    for name, scene, ps in avx_scenes(sc):
        f = ol.AvxFrame(scene, ps, threads)
        tot, main = [], []
        rc = 0
        for i in range(reps + 1):
            f.clear()
            rc, t_main, t_all = f.render()
            if i:
                tot.append(t_all * 1e3); main.append(t_main * 1e3)
        px = scene.width * scene.height
        rec = {"objects": len(ps), "triangles": scene.triangle_count * len(ps), "width": scene.width, "height": scene.height,
               "covered_pixels": f.covered(), "status": rc, "ms_median": float(np.median(tot)), "ms_best": float(min(tot)),
               "ms_submitting_thread": float(np.median(main)),
               "value": px / (float(np.median(tot)) * 1e-3) / 1e6}
        if renderer is not None:
            color = np.empty((scene.height, scene.width), np.uint32); z = np.empty((scene.height, scene.width), np.float32)
            g = []
            for i in range(reps + 1):
                color.fill(scene.clear_color); z.fill(scene.clear_depth)
                t0 = time.perf_counter()
                renderer.render_scene_host(scene, color, z, phong=True, object_ps=ps)
                if i:
                    g.append((time.perf_counter() - t0) * 1e3)
            rec["gpu_e2e_ms_median"] = float(np.median(g))
            rec["gpu_e2e_value"] = px / (float(np.median(g)) * 1e-3) / 1e6
            rec["gpu_covered_pixels"] = int((z != np.float32(scene.clear_depth)).sum())
        out["scenes"][name] = rec
    return out


def time_cpu(ol, s, threads, steps, warmup, kind, phong=False):
    """Time the reference's scalar path (verbatim build if present, else the port) on scene s."""
    lib_o = ol.oracle()
    pre = ol.oracle_render(s, phong=phong)        # untimed: tells which triangles crash the reference
    skip = pre["would_crash"]
    n = s.triangle_count
    os_ = ol.OracleScene(s)
    times = []
    if kind == "reference":
        lib = ol.ref()
        colors = [np.zeros((s.height, s.width), np.uint32) for _ in range(threads)]
        zs = [np.zeros((s.height, s.width), np.float32) for _ in range(threads)]
        bmps = (ol.RefLoadedBitmap * threads)(*[
            ol.RefLoadedBitmap(s.width, s.height, c.strides[0], c.ctypes.data) for c in colors])
        zptrs = (ol.f32p * threads)(*[z.ctypes.data_as(ol.f32p) for z in zs])
        cmd = os_.ref_commands(zs[0])
        ctx = ol.OrcFallbackCtx(os_.pos_p, os_.col_p, os_.nrm_p, os_.P, C.pointer(os_.orc), 1 if phong else 0,
                                os_.uvs_p if os_.orc_tex is not None else None,
                                C.pointer(os_.orc_tex) if os_.orc_tex is not None else None)
        fb = C.cast(lib_o.orc_ref_fallback, C.c_void_p)
        user = C.cast(C.pointer(ctx), C.c_void_p)
        lib.ref_set_texture(C.addressof(os_.ref_tex) if os_.ref_tex is not None else None)   # Object->Bitmap
        for i in range(warmup + steps):
            colors[0].fill(s.clear_color); zs[0].fill(s.clear_depth)
            t0 = time.perf_counter()
            lib.ref_render_triangles_mt(os_.pos_p, os_.col_p, os_.nrm_p, os_.uvs_p, n, os_.P,
                                        C.byref(cmd), bmps, zptrs, threads, skip.ctypes.data, fb, user, 1 if phong else 0)
            dt = time.perf_counter() - t0
            if i >= warmup:
                times.append(dt)
        lib.ref_set_texture(None)
        same = bool(np.array_equal(zs[0].view(np.uint32), pre["z"].view(np.uint32)) and
                    (np.array_equal(colors[0], pre["color"]) or pre["stats"]["TexelClamps"] > 0))
        note = (f"verbatim FillEdgeTable+DrawModel{' (PhongShading)' if phong else ''}{' (Bitmap)' if os_.ref_tex is not None else ''} per single-triangle object (oracle/_ref), {threads} threads with "
                f"private targets folded in submission order; {int(skip.sum())} of {n} triangles that null-deref "
                f"in the reference go through the oracle port; image identical to 1-thread oracle: {same}")
    else:
        assert not phong and os_.orc_tex is None, "the threaded port has no Phong / textured variant; the verbatim build is required"
        for i in range(warmup + steps):
            color, z, _ = ol.new_targets(s)
            t = ol._orc_target(color, z, None)
            st = ol.OrcStats()
            t0 = time.perf_counter()
            lib_o.orc_render_triangles_mt(os_.pos_p, os_.col_p, os_.nrm_p, n, os_.P, C.byref(os_.orc),
                                          C.byref(t), threads, C.byref(st))
            dt = time.perf_counter() - t0
            if i >= warmup:
                times.append(dt)
        note = f"oracle port (oracle/raster_oracle.c), {threads} threads"
    ms = 1e3 * float(np.mean(times))
    return {"ms_per_step": ms, "sample": f"{n} triangles of the frame per step, {len(times)} steps; {note}"}


# ------------------------------------------------------------------------------ our arm
PASSES = 5                     # timed passes of exactly K steps each; the line reports median and best


class Ctx:
    """What every leg of one bench process shares: rank geometry, device, stream, one renderer."""

    def __init__(self, args):
        import torch
        import torch.distributed as dist
        from cpu_renderer_b200 import api
        self.torch, self.dist, self.api, self.args = torch, dist, api, args
        self.rank = int(os.environ.get("RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        if not torch.cuda.is_available():
            raise SystemExit("bench.py needs a CUDA device: there is no CPU fallback for the product path")
        torch.cuda.set_device(self.local_rank)
        self.dev = torch.device("cuda", self.local_rank)
        if self.world > 1:
            dist.init_process_group("nccl", device_id=self.dev)
        self.stream = torch.cuda.Stream(device=self.dev)
        self.r = api.Renderer(self.local_rank)

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()

    def max_over_ranks(self, values):
        t = self.torch.tensor(list(values), dtype=self.torch.float64, device=self.dev)
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return [float(x) for x in t.tolist()]


def expected_hashes():
    p = os.path.join(ROOT, "tests", "golden", "bench_image_hashes.json")
    try:
        return json.load(open(p))
    except Exception:
        return {}


def run_leg(ctx, cfgname, K, Wm, *, passes=PASSES, scale=1.0, phong=False, textured=False, tile=None,
            e2e_steps=10, cpu_baseline=False, c1_whole_object=False, zbuffer_roofline=False):
    """One configuration, measured three ways: device-resident frames (CUDA events, `passes` passes of
    exactly K steps, max over ranks per pass), per-kernel durations, and end to end with host buffers."""
    torch, dist, api = ctx.torch, ctx.dist, ctx.api
    from cpu_renderer_b200 import imagehash, shard
    from cpu_renderer_b200 import scene as sc
    import copy
    rank, world, dev, stream, r, args = ctx.rank, ctx.world, ctx.dev, ctx.stream, ctx.r, ctx.args
    unit, workload = METRICS[cfgname]
    scene = build_scene(cfgname, rank, scale)
    if textured:
        scene = sc.textured(scene, TEX_SIZE, TEX_SIZE, lo=0.05, hi=0.95)
    ntri, W, H = scene.triangle_count, scene.width, scene.height
    wpad = (W + 63) // 64 * 64
    tile = tile or "0x0"                        # 0x0: the library picks (64x16 for small triangles, 128x8 otherwise)
    tw, th = (int(x) for x in tile.split("x"))
    r.set_stream(stream.cuda_stream)
    r.set_tile(tw, th)
    th = th or 32                               # row bands are split in multiples of this (any tile height divides it)
    flags = api.DEFER_VERDICT if args.defer_verdict else 0

    d_pos = torch.from_numpy(scene.positions).to(dev)
    d_col = torch.from_numpy(scene.colors).to(dev)
    d_nrm = torch.from_numpy(scene.normals).to(dev)
    mesh_uv, mesh_tex = None, None
    if textured:
        d_uv = torch.from_numpy(scene.uvs).to(dev)
        d_tex = torch.from_numpy(scene.texture.view(np.int32)).to(dev)
        dtex = api.device_texture(d_tex.data_ptr(), scene.texture.shape[1], scene.texture.shape[0], scene.texture.shape[1] * 4)
        mesh_uv, mesh_tex = d_uv.data_ptr(), C.pointer(dtex)

    # ---- the frames this rank renders in one step ------------------------------------------
    band_first, band_rows = 0, H
    mesh_flags = api.MESH_PHONG if phong else 0
    frames = []                                   # (device_mesh, game_render_commands, keepalive)
    if cfgname == "c4":
        band_first, band_rows = shard.band_rows(H, world, rank, th)
    if cfgname == "c5":
        for vi in shard.frame_range(C5_VIEWS, world, rank):
            P, D = c5_view(vi)
            sv = copy.copy(scene)
            sv.transform = copy.copy(scene.transform)
            sv.transform.distance_above_target = D
            cmd_v, keep_v = api.make_commands(sv)
            frames.append((api.device_mesh(d_pos.data_ptr(), d_col.data_ptr(), d_nrm.data_ptr(), ntri, api.v3(*P), mesh_flags,
                                           mesh_uv, mesh_tex), cmd_v, keep_v))
    else:
        cmd0, keep0 = api.make_commands(scene)
        frames.append((api.device_mesh(d_pos.data_ptr(), d_col.data_ptr(), d_nrm.data_ptr(), ntri,
                                       api.v3(*scene.object_p), mesh_flags, mesh_uv, mesh_tex), cmd0, keep0))
    # c1/c2/c3: one pre-cleared target pair per timed frame (clear outside the timed region).
    # c4/c5: one pair, cleared inside the step (a 16K^2 pair is 2 GiB; a real frame clears anyway).
    clear_in_step = cfgname in ("c4", "c5")
    nsets = 1 if clear_in_step else max(K, Wm, 1)
    colors = [torch.empty((band_rows, wpad), dtype=torch.int32, device=dev) for _ in range(nsets)]
    depths = [torch.empty((band_rows, wpad), dtype=torch.float32, device=dev) for _ in range(nsets)]
    targets = [api.device_target(c.data_ptr(), z.data_ptr(), W, H, wpad * 4, wpad, band_first, band_rows)
               for c, z in zip(colors, depths)]

    def clear_all():
        for c, z in zip(colors, depths):
            c.fill_(scene.clear_color); z.fill_(scene.clear_depth)
        torch.cuda.synchronize()

    def step(i):
        t = targets[i % nsets]
        for (m_, c_, _) in frames:
            if clear_in_step:
                r.clear_device(t, scene.clear_color, scene.clear_depth)
            r.render_device([m_], c_, t, flags)

    # ---- device-resident throughput: W warm-up steps, then `passes` passes of exactly K steps ----
    clear_all()
    with torch.cuda.stream(stream):
        for i in range(Wm):
            step(i)
        r.sync()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    pass_ms, launches = [], 0
    for _ in range(passes):
        clear_all()                               # every timed frame starts from cleared targets
        launches0 = r.stats()["KernelLaunches"]
        ctx.barrier(); torch.cuda.synchronize()
        with torch.cuda.stream(stream):
            ev0.record(stream)
            for i in range(K):
                step(i)
            ev1.record(stream)
            r.sync()
        torch.cuda.synchronize(); ctx.barrier()
        pass_ms.append(ev0.elapsed_time(ev1) / K)
        launches = r.stats()["KernelLaunches"] - launches0
    stats = r.stats()
    pass_ms = ctx.max_over_ranks(pass_ms)         # per pass: the slowest rank
    ms, ms_best = float(np.median(pass_ms)), float(min(pass_ms))

    # ---- the frame itself: hash of what rank 0 holds after one step (c4: gathered below) ------
    clear_all()
    with torch.cuda.stream(stream):
        step(0)
        r.sync()
    torch.cuda.synchronize()
    image_fnv = imagehash.image_fnv_torch(colors[0], W) if (world == 1 or cfgname != "c4") else None

    # ---- per-kernel durations (CUDA events on the launching stream, separate pass) ---------
    clear_all()
    r.set_profiling(True)
    stage = {k: [] for k in api.STAGES}
    t_end = time.time() + (1.0 if not clear_in_step else 0.3)
    with torch.cuda.stream(stream):
        i = 0
        while i < min(K, 8) or (time.time() < t_end and i < 64 * K):
            if i % nsets == 0 and i:
                clear_all()
            m_, c_, _ = frames[i % len(frames)]
            if clear_in_step:
                r.clear_device(targets[0], scene.clear_color, scene.clear_depth)
            r.render_device([m_], c_, targets[i % nsets])
            for k, v in r.stage_ms().items():
                stage[k].append(v)
            i += 1
    r.set_profiling(False)
    stage_ms = {k: float(np.median(v)) for k, v in stage.items()}
    dominant = max(stage_ms, key=stage_ms.get)

    # ---- the path's only exchange step: the finished colour images end up on rank 0 -------------
    # (a) fused: the raster kernel stores every finished tile a second time, straight into rank 0's
    #     peer-mapped image (b200r_set_gather_target, NVLink); only a stream-ordered barrier follows the frame
    # (b) for comparison: a torch.distributed gather (NCCL send/recv) behind every frame
    with_gather = None
    gather_nccl = None
    if world > 1:
        kg = min(K, 10)

        def timed_gather(after_step, passes_g):
            out = []
            for _ in range(passes_g):
                clear_all()
                ctx.barrier(); torch.cuda.synchronize()
                with torch.cuda.stream(stream):
                    ev0.record(stream)
                    for i in range(kg):
                        step(i)
                        after_step(i)                  # stream-ordered after this step's kernels
                    ev1.record(stream)
                    r.sync()
                torch.cuda.synchronize(); ctx.barrier()
                out.append(ev0.elapsed_time(ev1) / kg)
            return ctx.max_over_ranks(out)

        fused_ok = cfgname != "c5"                     # c5 renders 32 views per step into one target pair
        if fused_ok:
            slots = 1 if cfgname == "c4" else world
            fg = shard.FusedGather(r, api, H, W, wpad, slots, world, rank, dev, dst=0)
            fg.select(0 if cfgname == "c4" else rank)
            clear_all()
            with torch.cuda.stream(stream):
                for _ in range(2):                         # warm-up (peer mappings, communicator), untimed
                    step(0)
                    fg.finish()
                r.sync()
            torch.cuda.synchronize(); ctx.barrier()
            # what rank 0 now holds must be what the ranks rendered
            mine = imagehash.image_fnv_torch(colors[0], W) if cfgname != "c4" else None
            hashes = [None] * world
            dist.all_gather_object(hashes, mine)
            gather_ok = None
            if rank == 0:
                if cfgname == "c4":
                    image_fnv = imagehash.image_fnv_torch(fg.color[0], W)      # the image the bands assemble to
                else:
                    gather_ok = all(imagehash.image_fnv_torch(fg.color[q], W) == hashes[q] for q in range(world))
            g_ms = timed_gather(lambda i: fg.finish(), 3)
            fg.close()
            with_gather = {"ms_per_step": float(np.median(g_ms)), "ms_per_step_best": float(min(g_ms)),
                           "what": "fused: raster_kernel stores every finished tile into rank 0's peer-mapped image over NVLink "
                                   "(b200r_set_gather_target + CUDA IPC); a stream-ordered 4-byte all_reduce as the barrier",
                           "gather_bytes_per_step": int(H * wpad * 4 * (world - 1) // (world if cfgname == "c4" else 1)),
                           **({"images_match_the_ranks_frames": gather_ok} if gather_ok is not None else {})}
        if cfgname == "c4":
            gather = lambda t: shard.gather_bands(t, H, world, rank, th, dst=0)      # noqa: E731
        else:
            gl = [torch.empty_like(colors[0]) for _ in range(world)] if rank == 0 else None
            gather = lambda t: dist.gather(t, gl, dst=0)                             # noqa: E731
        gathered = None
        clear_all()
        with torch.cuda.stream(stream):
            for _ in range(2):                         # communicator set-up and warm-up, untimed
                step(0)
                gathered = gather(colors[0])
            r.sync()
        ctx.barrier(); torch.cuda.synchronize()
        if cfgname == "c4" and rank == 0 and not fused_ok:
            image_fnv = imagehash.image_fnv_torch(gathered, W)
        nccl_fnv = imagehash.image_fnv_torch(gathered, W) if (cfgname == "c4" and rank == 0) else None
        del gathered
        n_ms = timed_gather(lambda i: gather(colors[i % nsets]), 3 if not fused_ok else 1)
        gather_nccl = {"ms_per_step": float(np.median(n_ms)), "ms_per_step_best": float(min(n_ms)),
                       "what": "torch.distributed.gather (NCCL) of the finished colour image(s) to rank 0 after every step",
                       "gather_bytes_per_step": int(colors[0].numel() * 4 * (world - 1)),
                       **({"image_fnv": nccl_fnv} if nccl_fnv is not None else {})}
        if not fused_ok:
            with_gather, gather_nccl = gather_nccl, None

    whole_object = None
    shaded_vs_oracle = None
    # ---- end to end: host buffers in, host buffers out, copies inside the timed region ---------
    pin = lambda a: torch.from_numpy(a).pin_memory()              # noqa: E731
    e2e = None
    if e2e_steps > 0 and not clear_in_step:
        # c1/c2/c3: the reference-facing call b200r_render_objects (H2D vertices + targets, kernels, D2H targets)
        r.set_stream(0)
        e2e_steps = min(K, e2e_steps)
        hs = sc.Scene(scene.name, W, H, scene.transform, pin(scene.positions).numpy(), pin(scene.colors).numpy(),
                      pin(scene.normals).numpy(), pin(scene.uvs).numpy() if textured else scene.uvs,
                      scene.object_p, scene.ambient, scene.lights,
                      texture=pin(scene.texture.view(np.int32)).numpy().view(np.uint32) if textured else None)
        hcol = [torch.full((H, W), scene.clear_color, dtype=torch.int32).pin_memory().numpy().view(np.uint32)
                for _ in range(e2e_steps + 1)]
        hz = [torch.full((H, W), scene.clear_depth, dtype=torch.float32).pin_memory().numpy()
              for _ in range(e2e_steps + 1)]
        r.render_scene_host(hs, hcol[e2e_steps], hz[e2e_steps], phong=phong)       # warm-up (allocations)
        ctx.barrier(); torch.cuda.synchronize()
        t0 = time.perf_counter()
        for i in range(e2e_steps):
            r.render_scene_host(hs, hcol[i], hz[i], phong=phong)
        torch.cuda.synchronize()
        e2e_local = (time.perf_counter() - t0) / e2e_steps * 1e3
        covered = int((hz[0] != np.float32(scene.clear_depth)).sum())
        e2e_fnv = imagehash.image_fnv_numpy(hcol[0])
        e2e_api = "b200r_render_objects (host pointers, pinned)"
        shaded_vs_oracle = None
        if (phong or textured) and rank == 0 and world == 1 and scale * ntri <= 2_000_000:
            # the Phong / textured frames have no committed hash: compare the end-to-end frame with the CPU oracle's
            # here and say how many pixels differ and by how much (the Phong bar is +-1 LSB per channel)
            try:
                sys.path.insert(0, os.path.join(ROOT, "tests"))
                import oracle_lib as ol
                shaded_vs_oracle = compare_with_oracle(ol, scene, hcol[0], hz[0], phong)
            except Exception as e:
                shaded_vs_oracle = {"error": repr(e)}
        if c1_whole_object and rank == 0:
            # the same frame as ONE object through the whole-object mode, beside the
            # verbatim reference's own call pair on one host core (it is a single-threaded path)
            sys.path.insert(0, os.path.join(ROOT, "tests"))
            import oracle_lib as ol
            wo_c = [c.copy() for c in hcol[:3]]
            wo_z = [z_.copy() for z_ in hz[:3]]
            for c_, z_ in zip(wo_c, wo_z):
                c_.fill(scene.clear_color); z_.fill(scene.clear_depth)
            r.render_scene_host(hs, wo_c[2], wo_z[2], flags=api.WHOLE_OBJECT_AEL, phong=phong)   # warm-up
            t0 = time.perf_counter()
            for i in range(2):
                r.render_scene_host(hs, wo_c[i], wo_z[i], flags=api.WHOLE_OBJECT_AEL, phong=phong)
            wo_ms = (time.perf_counter() - t0) / 2 * 1e3
            whole_object = {"e2e_ms": wo_ms, "api": "b200r_render_objects + B200R_WHOLE_OBJECT_AEL",
                            "stopped_objects": r.stats()["StoppedObjects"]}
            if ol.ref_available():
                ref_t = []
                for i in range(3):
                    t0 = time.perf_counter()
                    ref = ol.ref_render_object(scene, phong=phong)
                    ref_t.append((time.perf_counter() - t0) * 1e3)
                whole_object["cpu_reference_ms"] = min(ref_t)
                whole_object["cpu_reference"] = "verbatim FillEdgeTable + DrawModel on the whole object, 1 thread, incl. clearing the targets"
                whole_object["depth_identical_to_reference"] = bool(np.array_equal(ref["z"].view(np.uint32), wo_z[0].view(np.uint32)))
                whole_object["colour_max_lsb_vs_reference"] = int(np.abs(ref["color"].view(np.uint8).astype(np.int16) -
                                                                         wo_c[0].view(np.uint8).astype(np.int16)).max())
        # a textured object's vertex colours are not uploaded (they never reach the image); its UVs and Bitmap are
        h2d_bytes = int(ntri * (96 if textured else 120) + 2 * W * H * 4 + (scene.texture.nbytes if textured else 0))
        d2h_bytes = int(2 * W * H * 4)
        del hcol, hz
    elif e2e_steps > 0:
        # c4/c5: vertices from pinned host memory every step, b200r_clear_device + b200r_render_device per
        # frame (band / view), colour and depth of every frame read back to pinned host memory.
        # c4 on N ranks: every rank uploads 1/N of the triangle list over PCIe and the ranks exchange their
        # slices over NVLink (all_gather) -- the host link carries the mesh once, not once per GPU.
        e2e_steps = min(K, e2e_steps, 3)
        sharded_upload = cfgname == "c4" and world > 1 and ntri % world == 0
        h_c = torch.empty((band_rows, wpad), dtype=torch.int32).pin_memory()
        h_z = torch.empty((band_rows, wpad), dtype=torch.float32).pin_memory()
        if sharded_upload:
            per = ntri // world * 3                                   # vertices per rank
            sl = slice(rank * per, (rank + 1) * per)
            h_pos, h_col, h_nrm = pin(scene.positions[sl]), pin(scene.colors[sl]), pin(scene.normals[sl])
        else:
            h_pos, h_col, h_nrm = pin(scene.positions), pin(scene.colors), pin(scene.normals)

        def e2e_once():
            with torch.cuda.stream(stream):
                if sharded_upload:
                    for d_all, h_part in ((d_pos, h_pos), (d_col, h_col), (d_nrm, h_nrm)):
                        d_all[sl].copy_(h_part, non_blocking=True)
                        dist.all_gather_into_tensor(d_all, d_all[sl])
                else:
                    d_pos.copy_(h_pos, non_blocking=True); d_col.copy_(h_col, non_blocking=True)
                    d_nrm.copy_(h_nrm, non_blocking=True)
                for (m_, c_, _) in frames:
                    r.clear_device(targets[0], scene.clear_color, scene.clear_depth)
                    r.render_device([m_], c_, targets[0])
                    h_c.copy_(colors[0], non_blocking=True); h_z.copy_(depths[0], non_blocking=True)
                r.sync()
            stream.synchronize()
        e2e_once()
        ctx.barrier(); torch.cuda.synchronize()
        t0 = time.perf_counter()
        for i in range(e2e_steps):
            e2e_once()
        torch.cuda.synchronize()
        e2e_local = (time.perf_counter() - t0) / e2e_steps * 1e3
        covered = int((h_z.numpy()[:, :W] != np.float32(scene.clear_depth)).sum())
        e2e_fnv = None
        e2e_api = ("pinned H2D of the vertex streams" +
                   (f" (1/{world} per rank + NCCL all_gather over NVLink)" if sharded_upload else "") +
                   " + b200r_clear_device/b200r_render_device per frame + D2H of colour and depth")
        h2d_bytes = int(ntri * 120 // (world if sharded_upload else 1))
        d2h_bytes = int(len(frames) * 2 * band_rows * wpad * 4)
        del h_c, h_z, h_pos, h_col, h_nrm
    if e2e_steps > 0:
        e2e_ms = ctx.max_over_ranks([e2e_local])[0]

    def to_value(ms_step, n_ranks):
        if cfgname == "c2":
            units = n_ranks * ntri                      # every rank: its own 1M-triangle frame
        elif cfgname in ("c1", "c3"):
            units = n_ranks * W * H
        elif cfgname == "c4":
            units = W * H                               # one frame, its bands spread over the ranks
        else:
            units = C5_VIEWS * ntri                     # 256 views in total, whatever the rank count
        return units / (ms_step * 1e-3) / 1e6

    peak, peak_src = measured_peak()
    nframes = len(frames)
    # algorithmic bytes of what THIS rank does in one step: read every vertex
    # attribute once per frame, load + store colour and depth once per pixel of its target
    # (textured: positions + normals + UVs = 96 B per triangle -- the vertex colours never reach the image)
    tri_bytes = 96.0 if textured else 120.0
    a_frame = nframes * (tri_bytes * ntri + 16.0 * W * band_rows)
    own_bytes = {"setup_kernel": tri_bytes * ntri, "raster_kernel": 16.0 * W * band_rows,
                 "tile_scan_kernel": 0.0, "scatter_kernel": 0.0}[dominant]
    achieved = own_bytes / (stage_ms[dominant] * 1e-3) / 1e9
    frame_achieved = a_frame / (ms * 1e-3) / 1e9
    traffic = ncu_traffic(cfgname, dominant)
    want = expected_hashes().get(cfgname, {}).get("color") if (scale == 1.0 and not phong and not textured) else None
    rec = {
        "metric": unit, "value": to_value(ms, world), "unit": unit, "n_gpus": world, "steps": K,
        "warmup": Wm, "ms_per_step": ms, "ms_per_step_best": ms_best, "ms_per_step_passes": pass_ms, "passes": passes,
        "higher_is_better": True, "scaling": SCALING[cfgname],
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload + (" -- per-pixel Phong shading" if phong else "") +
                               (f" -- textured, {TEX_SIZE}x{TEX_SIZE} ARGB8, perspective correct" if textured else ""),
                   "shading": ("phong" if phong else "gouraud") + ("+texture" if textured else ""), "triangles": ntri, "width": W, "height": H,
                   "parallelism": ({"c4": f"screen-space row bands x{world}", "c5": f"{C5_VIEWS} views over {world} ranks"}
                                   .get(cfgname, f"frame-parallel x{world}") if world > 1 else "1 GPU"),
                   "frames_per_step_per_rank": nframes, "band_rows": band_rows, "scale": scale,
                   "tile": (tile if tile != "0x0" else
                            "auto:" + ("64x16" if W * H / max(ntri, 1) < 4.0 else "128x8")),
                   "timing": f"median of {passes} passes of exactly {K} steps, per pass the max over ranks (CUDA events)",
                   "l2": "each timed frame streams >126 MB (vertices + records + pair lists + its own "
                         "pre-cleared target), i.e. inputs larger than L2; no explicit flush",
                   "targets": ("one target pair, b200r_clear_device inside every timed frame" if clear_in_step else
                               "one pre-cleared colour/depth pair per timed frame (clear outside the timed region)")},
        "frame_ms": ms / nframes,
        "gpu_launches": int(launches),
        **({"whole_object": whole_object} if whole_object else {}),
        **({"shaded_vs_oracle": shaded_vs_oracle} if (e2e_steps > 0 and not clear_in_step and shaded_vs_oracle) else {}),
        "stage_ms": stage_ms,
        "binner": {"binned_triangles": stats["Binned"], "segments": stats["Segments"], "spans": stats["Spans"],
                   "queue_entries": stats["TilePairs"], "tiles": stats["Tiles"], "reruns": stats["Reruns"]},
        "image_fnv": image_fnv,
        "image_fnv_oracle": want,
        "image_ok": (image_fnv == want) if (want and image_fnv) else None,
        "roofline": {"bound": "hbm", "kernel": dominant, "achieved": achieved, "peak": peak, "unit": "GB/s",
                     "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                     "algorithmic_bytes_per_launch": own_bytes,
                     "kernel_ms": stage_ms[dominant],
                     "kernel_share_of_step": stage_ms[dominant] / max(sum(stage_ms.values()), 1e-9),
                     "frame": {"algorithmic_bytes": a_frame, "achieved": frame_achieved,
                               "frac": frame_achieved / peak}},
    }
    if e2e_steps > 0:
        rec["e2e"] = {"value": to_value(e2e_ms, world), "unit": unit, "ms_per_step": e2e_ms,
                      "h2d_bytes_per_step": h2d_bytes, "d2h_bytes_per_step": d2h_bytes,
                      "steps": e2e_steps, "covered_pixels": covered, "api": e2e_api,
                      **({"image_fnv": e2e_fnv, "image_ok": (e2e_fnv == want) if want else None} if e2e_fnv else {})}
    if with_gather:
        with_gather["value"] = to_value(with_gather["ms_per_step"], world)
        rec["with_gather"] = with_gather
    if gather_nccl:
        gather_nccl["value"] = to_value(gather_nccl["ms_per_step"], world)
        rec["with_gather_nccl"] = gather_nccl
    if cpu_baseline and world == 1 and rank == 0:
        try:
            sys.path.insert(0, os.path.join(ROOT, "tests"))
            import oracle_lib as ol
            s, sample, threads = cpu_sample(cfgname, scene, sc)
            kind = "reference" if ol.ref_available() else "port"
            res = time_cpu(ol, s, threads, 3, 1, kind, phong)
            units = sample if unit == "Mtriangles/s" else W * H * (sample / ntri)
            rec["cpu_baseline"] = {"value": units / (res["ms_per_step"] * 1e-3) / 1e6, "unit": unit,
                                   "cores": threads, "kind": kind, "sample": res["sample"],
                                   "host": cpu_model(),
                                   "scalar_mt": {"value": units / (res["ms_per_step"] * 1e-3) / 1e6, "unit": unit, "cores": threads,
                                                 "whole_frame": sample == ntri,
                                                 "path": "verbatim FillEdgeTable + DrawModel per triangle, object-parallel over the host threads"}}
            try:
                # SURVEY.md 8d, CPU baseline (1): the same scalar path on ONE host thread (one step, no warm-up)
                one = time_cpu(ol, s, 1, 1, 0, kind, phong)
                rec["cpu_baseline"]["scalar_1t"] = {"value": units / (one["ms_per_step"] * 1e-3) / 1e6, "unit": unit, "cores": 1,
                                                    "ms_per_step": one["ms_per_step"], "whole_frame": sample == ntri,
                                                    "path": "verbatim FillEdgeTable + DrawModel per triangle, one host thread"}
            except Exception as e:
                rec["cpu_baseline"]["scalar_1t"] = {"error": repr(e)}
            try:
                r.set_stream(0); r.set_tile(64, 32)
                rec["cpu_baseline"]["avx_mt"] = avx_baseline(r)
            except Exception as e:
                rec["cpu_baseline"]["avx_mt"] = {"error": repr(e)}
        except Exception as e:  # the baseline is a reported number, never a reason to lose the line
            rec["cpu_baseline"] = {"value": None, "unit": unit, "cores": 0, "kind": "port",
                                   "sample": f"failed: {e!r}"}
    if (cpu_baseline or zbuffer_roofline) and world == 1 and rank == 0 and cfgname in ("c1", "c2", "c3"):
        # the reference algorithm's own z-buffer traffic and the raster kernel's issue roofline; both need the
        # frame's fragment counts, which the oracle counts on one host thread (c4 / c5 frames are too large for that)
        try:
            sys.path.insert(0, os.path.join(ROOT, "tests"))
            import oracle_lib as ol
            rec["roofline"]["equivalent_zbuffer"] = equivalent_zbuffer(ol, scene, tri_bytes, ms, peak)
            issue = issue_roofline(cfgname, rec["roofline"]["equivalent_zbuffer"]["fragments"], stage_ms["raster_kernel"], None)
            if issue:
                rec["roofline"]["issue"] = issue
        except Exception as e:
            rec["roofline"]["equivalent_zbuffer"] = {"error": repr(e)}
    # release this leg's device memory before the next one
    del colors, depths, targets, frames, d_pos, d_col, d_nrm
    torch.cuda.empty_cache()
    return rec


def leg_summary(rec):
    """The sub-record of a secondary leg inside the one JSON line."""
    keep = ("metric", "value", "unit", "ms_per_step", "ms_per_step_best", "passes", "steps", "scaling", "stage_ms",
            "gpu_launches", "image_fnv", "image_fnv_oracle", "image_ok", "with_gather", "with_gather_nccl", "binner")
    out = {k: rec[k] for k in keep if k in rec}
    out["config"] = {k: rec["config"][k] for k in ("workload", "triangles", "width", "height", "parallelism", "band_rows", "tile")}
    out["roofline"] = {k: rec["roofline"][k] for k in ("bound", "kernel", "achieved", "peak", "unit", "frac", "traffic",
                                                        "algorithmic_bytes_per_launch", "kernel_ms", "frame",
                                                        "equivalent_zbuffer", "issue") if k in rec["roofline"]}
    if "e2e" in rec:
        out["e2e"] = rec["e2e"]
    return out


def run_ours(args):
    ctx = Ctx(args)
    sampler = ClockSampler(ctx.local_rank) if ctx.rank == 0 else None
    K, Wm = args.steps, args.warmup
    main_leg = run_leg(ctx, args.config, K, Wm, scale=args.scale, phong=args.phong, textured=args.textured,
                       tile=args.tile, cpu_baseline=not args.no_cpu_baseline, c1_whole_object=args.config == "c1")
    legs = {}
    default_run = (args.config == "c2" and args.scale == 1.0 and not args.phong and not args.textured
                   and not args.tile and not args.no_legs)
    if default_run:
        # north_star's two numeric targets sit on C3 (4K fill, 1 GPU) and C4 (16K^2 row bands, every N):
        # short legs of both ride in the default line so that the driver's own runs carry them
        if ctx.world == 1:
            legs["c3"] = leg_summary(run_leg(ctx, "c3", min(K, 10), 3, passes=3, e2e_steps=5, zbuffer_roofline=True))
        legs["c4_bands"] = leg_summary(run_leg(ctx, "c4", min(K, 3), 3, passes=3, e2e_steps=2))
    clocks = sampler.stop() if sampler else None
    if ctx.rank == 0:
        line = main_leg
        if legs:
            line["legs"] = legs
        line["clocks"] = clocks
        emit_line(line)
    if ctx.world > 1:
        ctx.dist.destroy_process_group()


def main():
    claim_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--config", default="c2", choices=sorted(METRICS))
    ap.add_argument("--scale", type=float, default=1.0, help="shrink the triangle count (smoke runs only)")
    ap.add_argument("--textured", action="store_true", help="every object carries a 1024x1024 Bitmap and per-vertex UVs")
    ap.add_argument("--phong", action="store_true", help="per-pixel Phong shading (PhongShading = 1) instead of Gouraud")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--tile", default=None, help="WxH: 64x32 (default), 32x32, 128x16, 64x16")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-legs", action="store_true", help="default run without the short c3 / c4_bands legs")
    ap.add_argument("--defer-verdict", action="store_true",
                    help="device-resident frames with B200R_DEFER_VERDICT (no host wait for the binning verdict)")
    ap.add_argument("--ref-sample", type=int, default=0)
    ap.add_argument("--ref-threads", type=int, default=0)
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
