import sys; sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
import numpy as np, kat_scenes, oracle_lib as ol
from cpu_renderer_b200.api import Renderer
r = Renderer(0)
for name in ['horizontal','ties','slivers','lights_and_offset']:
    s = kat_scenes.all_scenes()[name]
    o = ol.oracle_render(s, with_prim=True, phong=True)
    color, z, _ = ol.new_targets(s)
    r.render_scene_host(s, color, z, phong=True)
    bad = np.argwhere(o['color'] != color)
    print(name, 'zdiff', (o['z'].view(np.uint32)!=z.view(np.uint32)).sum(), 'cdiff', len(bad))
    for (y,x) in bad[:6]:
        print('   y',y,'x',x,'oracle %08x gpu %08x prim'%(o['color'][y,x], color[y,x]), o['prim'][y,x])
