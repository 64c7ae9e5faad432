import sys, os, torch
ROOT=os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0,ROOT); sys.path.insert(0,os.path.join(ROOT,'ttdg-mgm_b200'))
import bench
from ttdg_b200 import synth
dev=torch.device('cuda',0)
for variant in ('pert','init'):
    m,opt=bench.build_ours(dev)
    if variant=='init':
        m.multi_matching_unsup.load_state_dict(synth.mgm_unsup_state(0))
        from ttdg_b200.optim import FlatSGD
        opt=FlatSGD(m.adapted_parameters(), lr=0.005, momentum=0.9, weight_decay=1e-4)
    inputs=[dict(d,image=d['image'].to(dev)) for d in bench.make_inputs(0)]
    w0=opt.flat_p.clone()
    for it in range(16):
        m.train()
        try:
            loss,_,_,_=m(inputs,branch='TTT')
        except Exception as e:
            print(variant,it,'EXC',e); break
        opt.zero_grad(); loss.backward(); gn=float(opt.flat_g.norm()); opt.step(1)
        aux=m.multi_matching_unsup.last_aux
        m.eval(); out=m(inputs)
        print(variant,it,'loss %.5f'%float(loss),'gnorm %.3e'%gn,'dW %.3e'%float((opt.flat_p-w0).norm()/w0.norm()),'sizes',aux['sizes'],'dets',[len(o['instances']) for o in out],'iters',int(aux['info'][0]))
