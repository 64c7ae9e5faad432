import sys, numpy as np, torch
import torch.nn.functional as F
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/ttdg-mgm_b200')
from oracle import detector_port as dp
from ttdg_b200 import synth, detector as det
from ttdg_b200.detector import _Subsample2
nhwc=lambda x: x.permute(0,2,3,1).contiguous()
nchw=lambda x: x.permute(0,3,1,2).contiguous()
sd=synth.detector_state_calibrated(0)
m=det.MaskRCNN(2).cuda(); m.load_state_dict(sd)
ims=[synth.fundus_like_image(100+i,64)['image'] for i in range(2)]
# ---- oracle with retained intermediates
sdg={k:v.clone().requires_grad_(k.endswith('weight') and 'norm' not in k and ('res3' in k or 'res4' in k or 'res5' in k or 'fpn' in k)) for k,v in sd.items()}
x=dp.preprocess(ims); p='backbone.bottom_up.'
with torch.no_grad():
    y=dp.conv_bn(x,sdg,p+'stem.conv1',stride=2,pad=3); y=F.max_pool2d(y,3,2,1)
    for b in range(3): y=dp.bottleneck(y,sdg,f'{p}res2.{b}.',1,b==0)
res={}; inter={}
for stage,blocks,stride in dp.STAGES[1:]:
    for b in range(blocks):
        y=dp.bottleneck(y,sdg,f'{p}{stage}.{b}.',stride if b==0 else 1,b==0); y.retain_grad(); inter[f'{stage}.{b}']=y
    res[stage]=y
lat=F.conv2d(res['res5'],sdg['backbone.fpn_lateral5.weight'],sdg['backbone.fpn_lateral5.bias']); lat.retain_grad()
p5=F.conv2d(lat,sdg['backbone.fpn_output5.weight'],sdg['backbone.fpn_output5.bias'],padding=1)
g=torch.Generator().manual_seed(9); w=torch.randn(p5.shape,generator=g)
(p5*w).sum().backward()
# ---- ours
bu=m.backbone.bottom_up
xin=det.preprocess(ims, torch.device('cuda'))
with torch.no_grad():
    yy=bu.stem.conv1(xin,relu=True)
    N,H,W,C=yy.shape; pp=torch.empty(N,(H-1)//2+1,(W-1)//2+1,C,device='cuda')
    from ttdg_b200 import _C; from ttdg_b200.ops import _p,_stream
    _C.check(_C.lib().ttdg_maxpool3x3s2(_p(yy),N,H,W,C,_p(pp),_stream()),'mp')
    yy=bu.res2(pp)
print('res2 fwd err', float((nchw(yy.cpu())-res.get('res2',y*0+0).shape and 0)) if False else '')
mine={}
for stage in ('res3','res4','res5'):
    for b,blk in enumerate(getattr(bu,stage)):
        yy=blk(yy); yy.retain_grad(); mine[f'{stage}.{b}']=yy
latc=m.backbone.fpn_lateral5(yy); latc.retain_grad()
p5c=m.backbone.fpn_output5(latc)
(p5c*nhwc(w).cuda()).sum().backward()
print('p5 fwd', float((nchw(p5c.detach().cpu())-p5.detach()).abs().max()/p5.detach().abs().max()))
print('lat grad', float((nchw(latc.grad.cpu())-lat.grad).abs().max()/lat.grad.abs().max()))
for k in reversed(list(inter)):
    a=nchw(mine[k].grad.cpu()); r=inter[k].grad
    f=float((nchw(mine[k].detach().cpu())-inter[k].detach()).abs().max()/inter[k].detach().abs().max())
    print(k,'grad err',float((a-r).abs().max()/r.abs().max()),'fwd err',f, 'frac relu-mask differs', float(((nchw(mine[k].detach().cpu())>0)!=(inter[k].detach()>0)).float().mean()))
print('---- res5 stage on the actual res4.5 output as a leaf')
xr=inter['res4.5'].detach().clone().requires_grad_(True)
sdg2={k:v.clone().requires_grad_(False) for k,v in sd.items()}
y=xr
outs=[]
for b in range(3): y=dp.bottleneck(y,sdg2,f'{p}res5.{b}.',2 if b==0 else 1,b==0)
g=torch.Generator().manual_seed(5); w2=torch.randn(y.shape,generator=g)
(y*w2).sum().backward()
xc=nhwc(xr.detach()).cuda().requires_grad_(True)
yc=xc
for blk in bu.res5: yc=blk(yc)
(yc*nhwc(w2).cuda()).sum().backward()
a=nchw(xc.grad.cpu()); r=xr.grad
print('dx err', float((a-r).abs().max()/r.abs().max()))
d=(a-r).abs()
print('err at even/even', float(d[:,:,::2,::2].max()), 'elsewhere', float(d[:,:,1::2,:].max()), float(d[:,:,:,1::2].max()), 'ref max', float(r.abs().max()))
print('ref nonzero at odd positions?', float(r[:,:,1::2,:].abs().max()), float(r[:,:,:,1::2].abs().max()), 'mine', float(a[:,:,1::2,:].abs().max()))
