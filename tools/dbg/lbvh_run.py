import sys, numpy as np, torch, time
sys.path.insert(0,'/root/repo')
from ray_tracing_b200 import host, scenes
sky=scenes.procedural_skybox(512)
text=scenes.synthetic_spheres_text(100000)
objs=host.parse_scene_string_large(text)
r=host.Renderer(num_gpus=1); r.upload_skybox(sky)
r.upload_scene(objs); r.synchronize()
W,H=1920,1080
frame=torch.zeros((H,W,3),dtype=torch.float32,device='cuda')
for kern,name in ((1,'pixel'),(2,'persistent'),(3,'wavefront')):
    for i in range(3):
        st=r.render_into(host.Camera(),frame.data_ptr(),W,H,stats=True,kernel=kern)
    print(name, st['render_ms'], st['rays'], st['rays']/st['render_ms']/1e3,'Mrays/s')
