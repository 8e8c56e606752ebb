import sys, numpy as np, torch
sys.path.insert(0,'/root/repo')
from ray_tracing_b200 import host, scenes
import bench
faces,_=bench.load_skybox_faces()
objs=host.parse_scene_string(scenes.builtin_scene_text(0))
r=host.Renderer(num_gpus=1); r.upload_skybox(faces); r.upload_scene(objs)
W,H=3840,2160
frame=torch.zeros((H,W,3),dtype=torch.float32,device='cuda')
cam=host.Camera()
for n in (1,2,4,8,16):
    for idx in sorted(set([0,n//2,n-1])):
        ts=[]
        for rep in range(5):
            st=r.render_into(cam,frame.data_ptr(),W,H,stats=True,interleave_count=n,interleave_index=idx)
            ts.append(st['render_ms'])
        print('n',n,'idx',idx,'ms %.3f'%min(ts),'rays',st['rays'], 'Grays/s %.2f'%(st['rays']/min(ts)/1e6))
# contiguous bands of 1/8
for k in range(8):
    st=r.render_into(cam,frame.data_ptr(),W,H,stats=True,rows=(k*270,(k+1)*270))
    print('band',k,'ms %.3f'%st['render_ms'],'rays',st['rays'])
