import numpy as np, sys
sys.path.insert(0,'/root/repo')
from ray_tracing_b200 import host, scenes
from oracle.bindings import Port, procedural_skybox
sky=procedural_skybox(64,seed=7)
objs=host.parse_scene_string(scenes.builtin_scene_text(0))
port=Port(); want,_=port.render(port.world(objs,sky),640,360,1,1,0)
def bits(a): return np.ascontiguousarray(a,np.float32).view(np.uint32)
for n in (1,2,1,2):
    r=host.Renderer(num_gpus=n); r.upload_skybox(sky); r.upload_scene(objs)
    got,st=r.render_frame(host.Camera(),640,360,1)
    d=(bits(got)!=bits(want))
    print('ngpu',n,'mismatch px',d.any(axis=-1).sum(), 'rows with mismatch', np.unique(np.nonzero(d.any(axis=(1,2)))[0])[:10], st['rays'])
    r.close()
