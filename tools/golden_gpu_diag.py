import sys, os, glob, numpy as np
ROOT='/root/repo'; sys.path[:0]=[ROOT, ROOT+'/cat-generator_b200', ROOT+'/tests']
from oracle import pyoracle as po
from catgen import lib, models
import test_golden as tg
L=lib.load(); lib.init(0)
for name in tg.CASES:
    if not name.startswith('D32'): continue
    fx,net,m=tg.load(name); C,B=int(fx['C']),int(fx['B'])
    for eng in (0,1):
        lib.check(L.cg_set_conv_engine(eng))
        d=models.create_D((C,32,32),True); d.set_params(m.params)
        if int(fx['train']): d.training(); d.set_masks(fx['masks'],B,1)
        else: d.evaluate()
        out,pre=d.forward(fx['inp'],with_pre=True); d.zeroGradParameters(); gi=d.backward(fx['inp'],fx['gout']); gp=d.get_grads()
        print(name,'engine',eng,'out',np.abs(out[:,0]-fx['out']).max(),'pre',np.abs(pre-fx['pre']).max(),'gi l2',tg.l2rel(gi,fx['ginp']),'gi max',tg.rel(gi,fx['ginp']),'gp',tg.rel(gp[fx['gparam_idx']],fx['gparam_sample'],float(fx['gparam_max'])),flush=True)
