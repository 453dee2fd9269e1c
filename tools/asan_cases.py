import sys; import os; R=os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0,R); sys.path.insert(0,os.path.join(R,'tests'))
import numpy as np, torch
import harness
from admm_deconv_b200 import _lib
from cases import make_case
from parity import check_forward, check_backward
be=harness.EmuBackend(_lib.AdmmTvLib(sys.argv[1]))
for iso in (False, True):
    for (M,N,P,B,kh,kw,K) in [(32,32,1,2,0,0,3),(32,64,3,1,5,4,3),(64,32,1,3,3,3,3),(128,32,1,2,3,3,2),(32,512,1,1,3,3,2),(512,32,1,2,3,3,2)]:
        y,h,_=make_case(M,N,P,B,kh,kw,50+M+K)
        xbar=torch.from_numpy(np.random.default_rng(K).standard_normal((M,N,P,B)))
        print(iso,M,N,P,B,kh,kw,K, check_backward(be,y,h,0.05,0.3,iso,K,xbar,act='relu1',bias=0.01,tol=1e-5,tol_scalar=2e-4), flush=True)
y,_,_=make_case(32,32,3,2,0,0,77)
x=be.forward_grouped(y.numpy(),[0.03]*5,[0.05,0.1,0.2,0.4,0.8],None,True,3,groups=5,shared_input=True,concat=True,act='relu1')
ys=[make_case(32,32,1,1,3,5,900+b)[0] for b in range(3)]
import torch
h=torch.cat([make_case(32,32,1,1,3,5,900+b)[1][:,:,:,0] for b in range(3)],dim=2)
x=be.forward_grouped(torch.cat(ys,dim=3).numpy(),[0.02,0.04,0.06],[0.2,0.3,0.4],h.numpy(),False,3,groups=3)
print('grouped ok')
