import sys, os, torch
import torch.nn.functional as F
ROOT=os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0,ROOT); sys.path.insert(0,os.path.join(ROOT,'ttdg-mgm_b200'))
from ttdg_b200 import detector as det, synth
from oracle import detector_port as dp
nhwc=lambda x: x.permute(0,2,3,1).contiguous(); nchw=lambda x: x.permute(0,3,1,2).contiguous()
sd=synth.detector_state_calibrated(0)
m=det.MaskRCNN(2).cuda(); m.load_state_dict(sd)
stage,blocks,cin,H='res5',3,1024,4
q=f'backbone.bottom_up.{stage}.'; seq=getattr(m.backbone.bottom_up,stage)
g=torch.Generator().manual_seed(1)
x=torch.randn(2,cin,H,H,generator=g).requires_grad_(True)
sdg={k:v.clone().requires_grad_(k.startswith(q) and k.endswith('weight') and 'norm' not in k) for k,v in sd.items()}
y=x; inter=[]
for b in range(blocks):
    y=dp.bottleneck(y,sdg,f'{q}{b}.',2 if b==0 else 1,b==0); y.retain_grad(); inter.append(y)
w=torch.randn(y.shape,generator=g); (y*w).sum().backward()
for mode in ('simt','tf32x3'):
    det.set_conv_mode(mode)
    for p in m.parameters(): p.grad=None
    xc=nhwc(x.detach()).cuda().requires_grad_(True); yc=xc; mine=[]
    for b in range(blocks):
        yc=seq[b](yc); yc.retain_grad(); mine.append(yc)
    (yc*nhwc(w).cuda()).sum().backward()
    e=lambda a,r: '%.2e/%.2e'%(float((a-r).abs().max()/r.abs().max()), float((a-r).norm()/r.norm()))
    print(mode,'dx',e(nchw(xc.grad.cpu()),x.grad))
    for b in reversed(range(blocks)):
        print('  block',b,'out-grad',e(nchw(mine[b].grad.cpu()),inter[b].grad),'fwd',e(nchw(mine[b].detach().cpu()),inter[b].detach()),'maskdiff',int(((nchw(mine[b].detach().cpu())>0)!=(inter[b].detach()>0)).sum()),
              ' '.join(f"{n}:{e(getattr(seq[b],n).weight.grad.permute(3,2,0,1).cpu(), sdg[f'{q}{b}.{n}.weight'].grad)}" for n in ('conv3','conv2','conv1')))
