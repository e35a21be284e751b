import sys, os, ctypes as C
sys.path.insert(0,'.'); sys.path.insert(0,'oracle'); sys.path.insert(0,'tests')
import numpy as np, torch, octvr_b200 as vr, oracle as O, util
cfg,width,in_size=util.named_rig('rig6'); n=6; iw,ih=in_size
ot=O.build_template(cfg,width)
for d in ot.inputs: d['vignette']=None
t=vr.MapperTemplate.from_arrays(ot.out_size, ot.inputs, ot.seam_masks)
m=vr.Mapper(t,[in_size]*n,blend=-1,enable_gain_compensator=True)
W,H=t.out_size
ins=[]
for c in range(n):
    y,u,v=util.i420_planes(util.noise_frame(c,iw,ih),iw,ih)
    ins.append(torch.from_numpy(np.concatenate([y,np.concatenate([u,v],1)],0)).cuda())
out=torch.zeros((H*3//2,W),dtype=torch.uint8,device='cuda')
for k in range(5):
    m.stitch_packed(ins,out); torch.cuda.synchronize()
    d=(C.c_ulonglong*5)(); vr.lib().octvr_mapper_debug_gain_ns(m._h, d); d=list(d)
    print("last CTA: stats %.1f us, reduce %.1f, solve %.1f, tables %.1f" % ((d[1]-d[0])/1e3,(d[2]-d[1])/1e3,(d[3]-d[2])/1e3,(d[4]-d[3])/1e3))
