import sys, numpy as np
sys.path.insert(0,'/root/repo')
from ray_tracing_b200 import host, scenes
from oracle.bindings import procedural_skybox
sky=procedural_skybox(64,seed=7)
r=host.Renderer(num_gpus=1); r.upload_skybox(sky)
for k in (0,1,2):
    r.upload_scene(host.parse_scene_string(scenes.builtin_scene_text(k)))
    for kern in (1,2):
        f,st=r.render_frame(host.Camera(),200,120,1,kernel=kern); f,st=r.render_frame(host.Camera(),203,121,4,num_columns=3,kernel=kern)
    r.render_sweep(host.Camera(),192,108,16)
objs=host.parse_scene_string_large(scenes.synthetic_spheres_text(3000))
r.upload_scene(objs); f,st=r.render_frame(host.Camera(),160,90,1); print('lbvh rays',st['rays'])
print(r.div_check(1,8,64)); r.close(); print('done')
