import numpy as np, sys
sys.path.insert(0,'/root/repo')
from ray_tracing_b200 import host, scenes
from oracle.bindings import Port, procedural_skybox
sky=procedural_skybox(64,seed=7)
objs=host.parse_scene_string(scenes.builtin_scene_text(0))
port=Port(); W,H,s=640,360,1
ref,_=port.render(port.world(objs,sky),W,H,s,1,0)
def bits(a): return np.ascontiguousarray(a,np.float32).view(np.uint32)
def rep(name,a):
    d=(bits(a)!=bits(ref)).any(axis=-1); print(name,'mismatch px',d.sum(), 'rows', np.unique(np.nonzero(d)[0])[:6])
r=host.Renderer(num_gpus=1); r.upload_skybox(sky); r.upload_scene(objs)
want,st1=r.render_frame(host.Camera(),W,H,s); rep('want(before sweep)',want)
sweep1,_=r.render_sweep(host.Camera(),W,H,8); rep('want(after sweep)',want)
w2,_=r.render_frame(host.Camera(),W,H,s); rep('render after sweep',w2)
r.close()
