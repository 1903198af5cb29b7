#!/usr/bin/env python
"""Test infrastructure: renders the bench configurations with the CPU ORACLE (oracle/liboracle.so) and
writes the image hashes bench.py compares its GPU frames with (tests/golden/bench_image_hashes.json).

    python tools/make_image_hashes.py [c2 c3 c4 c5]

C4 at full size (20 M triangles, 16384^2) needs a 2 GiB target pair per oracle thread."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
OUT = os.path.join(ROOT, "tests", "golden", "bench_image_hashes.json")


def main():
    import oracle_lib as ol
    from cpu_renderer_b200 import imagehash as ih
    from cpu_renderer_b200 import scene as sc
    import bench
    todo = sys.argv[1:] or ["c2", "c3", "c4", "c5"]
    res = json.load(open(OUT)) if os.path.exists(OUT) else {}
    for cfg in todo:
        t0 = time.time()
        if cfg == "c5":
            mesh = sc.construct_sphere(sc.C5_STEP_COUNT)
            for view in (0, 255):
                s = sc.c5_scene(mesh, view)
                w = ol.oracle_render(s, threads=8)
                res[f"c5_view{view}"] = {"color": ih.image_fnv_numpy(w["color"]), "depth": ih.image_fnv_numpy(w["z"]),
                                         "fragments": int(w["stats"]["Fragments"])}
        else:
            s = bench.build_scene(cfg, 0, 1.0)
            w = ol.oracle_render(s, threads=3 if cfg == "c4" else 8)
            res[cfg] = {"color": ih.image_fnv_numpy(w["color"]), "depth": ih.image_fnv_numpy(w["z"]),
                        "fragments": int(w["stats"]["Fragments"]), "depth_passes": int(w["stats"]["DepthPasses"]),
                        "triangles": s.triangle_count, "width": s.width, "height": s.height}
        print(cfg, res.get(cfg), f"{time.time() - t0:.1f}s", flush=True)
        json.dump(res, open(OUT, "w"), indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
