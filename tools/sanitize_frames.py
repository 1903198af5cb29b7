#!/usr/bin/env python
"""Developer tool: the frames compute-sanitizer is run on (profiles/r02_sanitizer_*.txt): smoke()'s two
frames plus one per-pixel Phong and one textured + Phong frame, through the host-pointer C ABI, each checked
against the CPU oracle.  No torch import: the process under the sanitizer holds only our library's kernels."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    import oracle_lib as ol
    from cpu_renderer_b200 import scene as sc
    from cpu_renderer_b200.api import Renderer
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 4000
    r = Renderer(0)
    base = sc.triangle_soup("san", 0x5A, n, 640, 360, 2.0, 40.0)
    for name, s, phong in (("gouraud", base, False), ("phong", base, True),
                           ("textured_phong", sc.textured(base, 64, 48), True)):
        for tile in ((64, 32), (128, 16)):
            want = ol.oracle_render(s, phong=phong)
            color, z, _ = ol.new_targets(s)
            r.set_tile(*tile)
            r.render_scene_host(s, color, z, phong=phong)
            zdiff = int((want["z"].view(np.uint32) != z.view(np.uint32)).sum())
            lsb = int(np.abs(want["color"].view(np.uint8).astype(np.int16) - color.view(np.uint8).astype(np.int16)).max())
            print(f"{name} {tile}: zdiff={zdiff} colour max LSB={lsb}", flush=True)
            assert zdiff == 0 and lsb <= 1
    r.close()


if __name__ == "__main__":
    main()
