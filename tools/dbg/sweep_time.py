import sys, numpy as np, torch, time
sys.path.insert(0,'/root/repo')
from ray_tracing_b200 import host, scenes
import bench
faces,_=bench.load_skybox_faces()
r=host.Renderer(num_gpus=1); r.upload_skybox(faces)
r.upload_scene(host.parse_scene_string(scenes.builtin_scene_text(0)))
W,H=1920,1080
frame=torch.zeros((H,W,3),dtype=torch.float32,device='cuda')
cam=host.Camera()
for rep in range(3):
    _,st=r.render_sweep(cam,W,H,16,0,ptr=frame.data_ptr(),stats=True)
print('with stats: device sum ms', st['render_ms'], 'launches', st['kernel_launches'])
for rep in range(5):
    torch.cuda.synchronize(); t0=time.perf_counter()
    r.render_sweep(cam,W,H,16,0,ptr=frame.data_ptr(),stats=False); r.synchronize()
    t1=time.perf_counter()
print('wall ms', (t1-t0)*1e3)
# accumulate pass cost by scale
for s in (16,8,4,2,1):
    r.accum_reset()
    for rep in range(3):
        st=r.render_into(cam,frame.data_ptr(),W,H,stats=True,scale=s,accumulate=1)
    st2=r.render_into(cam,frame.data_ptr(),W,H,stats=True,scale=s,accumulate=0)
    print('scale',s,'accumulate ms %.3f'%st['render_ms'],'plain ms %.3f'%st2['render_ms'])
