import sys, numpy as np
sys.path.insert(0,'/root/repo')
from ray_tracing_b200 import host, scenes
import bench
faces,_=bench.load_skybox_faces()
r=host.Renderer(num_gpus=1); r.upload_skybox(faces)
for sc in (0,1,2):
    r.upload_scene(host.parse_scene_string(scenes.builtin_scene_text(sc)))
    for (W,H) in ((1280,720),(3840,2160)):
        e,se=r.render_frame(host.Camera(),W,H,1,variant=0)
        f,sf=r.render_frame(host.Camera(),W,H,1,variant=1)
        qa=host.quantize_frame(e).astype(np.int32); qb=host.quantize_frame(f).astype(np.int32)
        d=np.abs(qa-qb).max(axis=-1)
        print('scene',sc,W,H,'within 1 LSB: %.5f%%'%(100*(d<=1).mean()),'identical 8-bit: %.3f%%'%(100*(d==0).mean()),'rays',se['rays'],sf['rays'],'ms exact %.3f fast %.3f'%(se['render_ms'],sf['render_ms']))
