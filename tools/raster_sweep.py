#!/usr/bin/env python
"""Developer tool (not part of the product): per-kernel times of one config for a sweep of the
raster kernel's run-time thresholds and tile shapes, device-resident frames, CUDA events.

    python tools/raster_sweep.py --config c3 --tiles 128x16,256x8 --pend 4,8,12 --refill 8 [--phong] [--textured]
    B200R_LIB=cpu_renderer_b200/libb200raster_r4.so python tools/raster_sweep.py ...   (a build variant)
"""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="c3")
    ap.add_argument("--scale", type=float, default=1.0)
    ap.add_argument("--tiles", default="128x16")
    ap.add_argument("--pend", default="4")
    ap.add_argument("--refill", default="8")
    ap.add_argument("--frames", type=int, default=20)
    ap.add_argument("--phong", action="store_true")
    ap.add_argument("--textured", action="store_true")
    ap.add_argument("--band", default="0/1", help="i/n: render only band i of n row bands (what rank i of n GPUs does)")
    args = ap.parse_args()
    import torch
    import bench
    from cpu_renderer_b200 import api, scene as sc
    dev = torch.device("cuda", 0)
    scene = bench.build_scene(args.config, 0, args.scale)
    if args.textured:
        scene = sc.textured(scene, bench.TEX_SIZE, bench.TEX_SIZE, lo=0.05, hi=0.95)
    W, H, ntri = scene.width, scene.height, scene.triangle_count
    wpad = (W + 63) // 64 * 64
    d_pos = torch.from_numpy(scene.positions).to(dev)
    d_col = torch.from_numpy(scene.colors).to(dev)
    d_nrm = torch.from_numpy(scene.normals).to(dev)
    uv, tex = None, None
    if args.textured:
        import ctypes as C
        d_uv = torch.from_numpy(scene.uvs).to(dev)
        d_tex = torch.from_numpy(scene.texture.view(np.int32)).to(dev)
        dtex = api.device_texture(d_tex.data_ptr(), scene.texture.shape[1], scene.texture.shape[0], scene.texture.shape[1] * 4)
        uv, tex = d_uv.data_ptr(), C.pointer(dtex)
    from cpu_renderer_b200 import shard
    bi, bn = (int(x) for x in args.band.split("/"))
    first, rows = shard.band_rows(H, bn, bi, 32)
    color = torch.empty((rows, wpad), dtype=torch.int32, device=dev)
    depth = torch.empty((rows, wpad), dtype=torch.float32, device=dev)
    target = api.device_target(color.data_ptr(), depth.data_ptr(), W, H, wpad * 4, wpad, first, rows)
    cmd, keep = api.make_commands(scene)
    mesh = api.device_mesh(d_pos.data_ptr(), d_col.data_ptr(), d_nrm.data_ptr(), ntri, api.v3(*scene.object_p),
                           api.MESH_PHONG if args.phong else 0, uv, tex)
    ref_hash = None
    for tile in args.tiles.split(","):
        tw, th = (int(x) for x in tile.split("x"))
        for pend in args.pend.split(","):
            for refill in args.refill.split(","):
                os.environ["B200R_PEND"] = pend
                os.environ["B200R_REFILL"] = refill
                r = api.Renderer(0)
                r.set_tile(tw, th)
                r.set_profiling(True)
                stage = {k: [] for k in api.STAGES}
                for i in range(args.frames + 3):
                    color.fill_(scene.clear_color); depth.fill_(scene.clear_depth)
                    torch.cuda.synchronize()
                    r.render_device([mesh], cmd, target)
                    ms = r.stage_ms()
                    if i >= 3:
                        for k, v in ms.items():
                            stage[k].append(v)
                h = int(torch.sum(color.view(-1).to(torch.int64) * 2654435761 % 4294967291).item()) ^ \
                    int(torch.sum(depth.view(torch.int32).view(-1).to(torch.int64) % 1000003).item())
                if ref_hash is None:
                    ref_hash = h
                out = {k: round(float(np.median(v)), 4) for k, v in stage.items()}
                out["frame"] = round(sum(out.values()), 4)
                print(json.dumps({"lib": os.path.basename(api.LIB_PATH), "config": args.config, "tile": tile, "pend": int(pend),
                                  "refill": int(refill), **out, "same_image": h == ref_hash}), flush=True)
                r.close()


if __name__ == "__main__":
    main()
