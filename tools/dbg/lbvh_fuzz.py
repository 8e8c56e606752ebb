import sys, numpy as np, torch, time, ctypes
sys.path.insert(0,'/root/repo')
from ray_tracing_b200 import host, scenes
sky=scenes.procedural_skybox(512)
objs=host.parse_scene_string_large(scenes.synthetic_spheres_text(100000))
r=host.Renderer(num_gpus=1); r.upload_skybox(sky)
L=host.load_library(); L.rt_lbvh_debug_set.argtypes=[ctypes.c_double,ctypes.c_double]
W,H=1920,1080
frame=torch.zeros((H,W,3),dtype=torch.float32,device='cuda')
ref=None
for k,sl in ((32,1e-3),(16,1e-3),(8,1e-3),(0,1e-3),(32,1e-4),(32,0.0),(0,0.0)):
    L.rt_lbvh_debug_set(k,sl); r.upload_scene(objs)
    for i in range(3): st=r.render_into(host.Camera(),frame.data_ptr(),W,H,stats=True)
    f=frame.cpu().numpy().copy()
    if ref is None: ref=f
    print('k',k,'slack',sl,'ms %.2f'%st['render_ms'],'rays',st['rays'],'pixels differing from k=32:',int((f!=ref).any(axis=-1).sum()))
