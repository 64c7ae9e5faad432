import sys, os, torch, time
ROOT=os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0,ROOT); sys.path.insert(0,os.path.join(ROOT,'ttdg-mgm_b200'))
import bench
from ttdg_b200 import detector as D
dev=torch.device('cuda',0)
m,opt=bench.build_ours(dev); det=m._det[0]
images=[d['image'].to(dev) for d in bench.make_inputs(0)]
with torch.no_grad():
    feats=det.backbone(D.preprocess(images,dev))
rpn=det.proposal_generator
import torch.autograd.profiler as prof
for training in (True,False):
    for _ in range(2): rpn(feats,(512,512),training)
    torch.cuda.synchronize(); t=time.perf_counter()
    for _ in range(3): rpn(feats,(512,512),training)
    torch.cuda.synchronize(); print('training',training,'ms per call',(time.perf_counter()-t)/3*1e3)
with torch.profiler.profile(activities=[torch.profiler.ProfilerActivity.CPU, torch.profiler.ProfilerActivity.CUDA]) as p:
    rpn(feats,(512,512),True); torch.cuda.synchronize()
print(p.key_averages().table(sort_by='cuda_time_total', row_limit=14, max_name_column_width=60))
