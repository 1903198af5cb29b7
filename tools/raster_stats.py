#!/usr/bin/env python
"""Developer tool: what the raster kernel's lanes do, counted by a -DB200R_STATS build of the library.

    tools/build_variant.sh stats -DB200R_STATS
    B200R_LIB=cpu_renderer_b200/libb200raster_stats.so python tools/raster_stats.py --config c3 [--scale 0.01] [--tile 128x16]
"""
import argparse
import ctypes as C
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

NAMES = {0: "rounds", 1: "tested_px", 2: "replay_px", 3: "entries", 4: "culled", 5: "no_overlap", 6: "update_exec",
         7: "update_lanes", 8: "cas", 9: "cas_retry", 10: "exact_fail", 11: "refill_exec", 12: "refill_lanes",
         13: "clk_stage", 14: "clk_expand_floor", 15: "clk_raster", 16: "clk_tail_wait", 24: "clk_writeback", 25: "clk_samples",
         17: "tiles", 18: "tiles_empty", 19: "busy_lanes", 20: "pend_lanes", 21: "need_lanes", 22: "loop_iters", 23: "round_px_slots"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="c3")
    ap.add_argument("--scale", type=float, default=1.0)
    ap.add_argument("--tile", default="128x16")
    args = ap.parse_args()
    import torch
    import bench
    from cpu_renderer_b200 import api
    dev = torch.device("cuda", 0)
    scene = bench.build_scene(args.config, 0, args.scale)
    W, H, ntri = scene.width, scene.height, scene.triangle_count
    wpad = (W + 63) // 64 * 64
    d_pos = torch.from_numpy(scene.positions).to(dev)
    d_col = torch.from_numpy(scene.colors).to(dev)
    d_nrm = torch.from_numpy(scene.normals).to(dev)
    color = torch.empty((H, wpad), dtype=torch.int32, device=dev)
    depth = torch.empty((H, wpad), dtype=torch.float32, device=dev)
    target = api.device_target(color.data_ptr(), depth.data_ptr(), W, H, wpad * 4, wpad, 0, H)
    cmd, keep = api.make_commands(scene)
    mesh = api.device_mesh(d_pos.data_ptr(), d_col.data_ptr(), d_nrm.data_ptr(), ntri, api.v3(*scene.object_p), 0, None, None)
    tw, th = (int(x) for x in args.tile.split("x"))
    r = api.Renderer(0)
    r.set_tile(tw, th)
    r.set_profiling(True)
    lib = C.CDLL(api.LIB_PATH)
    buf = (C.c_ulonglong * 32)()
    for i in range(3):
        color.fill_(scene.clear_color); depth.fill_(scene.clear_depth)
        torch.cuda.synchronize()
        lib.b200r_debug_raster_stats(buf, 1)
        r.render_device([mesh], cmd, target)
        ms = r.stage_ms()
    assert lib.b200r_debug_raster_stats(buf, 1) == 0
    st = {NAMES.get(i, str(i)): int(buf[i]) for i in range(32) if buf[i]}
    st["stage_ms"] = ms
    st["config"] = args.config; st["scale"] = args.scale; st["tile"] = args.tile
    print(json.dumps(st))
    n = max(st.get("clk_samples", 1), 1)
    print("per tile-warp cycles: stage %.0f expand+floor %.0f raster %.0f tail-wait %.0f writeback %.0f" % tuple(
        st.get(k, 0) / n for k in ("clk_stage", "clk_expand_floor", "clk_raster", "clk_tail_wait", "clk_writeback")))
    rd = max(st.get("rounds", 1), 1)
    print("per round: px slots used %.1f/%d, busy %.1f pend %.1f need %.1f lanes; update every %.2f rounds with %.1f lanes" % (
        st.get("round_px_slots", 0) / rd, 32 * 4, st.get("busy_lanes", 0) / rd, st.get("pend_lanes", 0) / rd, st.get("need_lanes", 0) / rd,
        rd / max(st.get("update_exec", 1), 1), st.get("update_lanes", 0) / max(st.get("update_exec", 1), 1)))
    e = max(st.get("entries", 1), 1)
    print("entries %d culled %.1f%% replay px/entry %.1f tested px/entry %.1f; cas %d retry %d exact_fail %d" % (
        e, 100.0 * st.get("culled", 0) / e, st.get("replay_px", 0) / e, st.get("tested_px", 0) / e, st.get("cas", 0), st.get("cas_retry", 0),
        st.get("exact_fail", 0)))


if __name__ == "__main__":
    main()
