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

# sizes without a register-FFT plan (generic kernels), mixed dispatch (generic dim-1 + tuned dim-2), 3-smooth length
for iso in (False, True):
    for (M,N,P,B,kh,kw,K) in [(20,24,1,2,3,3,3),(33,17,3,1,5,4,3),(7,5,1,1,0,0,2),(48,32,1,2,3,3,3),(96,32,1,1,3,3,2)]:
        y,h,_=make_case(M,N,P,B,kh,kw,60+M+K)
        xbar=torch.from_numpy(np.random.default_rng(K).standard_normal((M,N,P,B)))
        print('generic',iso,M,N,P,B,kh,kw,K, check_backward(be,y,h,0.05,0.3,iso,K,xbar,act='relu1',bias=0.01,tol=1e-5,tol_scalar=5e-4), flush=True)
# losses and batch assembly (include/admmtv_loss.h, admmtv_batch.h): ragged tiles
import test_emu_losses as TL
lib=be.lib
for shp in [(70,37,1,2),(16,12,3,2),(5,3,2,1)]:
    x,yy=TL._images(*shp,3); print('gmsd',shp,TL.run_gmsd(lib,x,yy)[0], flush=True)
for shp in [(45,70,1,2),(24,20,3,2),(11,11,1,1)]:
    x,yy=TL._images(*shp,3); print('ssim',shp,TL.run_ssim(lib,x,yy)[0], TL.run_ssim(lib,x,yy,taps=[0.2]*5)[0], flush=True)
# round 2c: streaming GMSD strips (ragged rows / columns, planes smaller than a strip), register-blocked SSIM, arbitrary windows, pad_symmetric
for shp in [(31,65,1,1),(61,129,2,1),(3,2,1,1),(64,64,1,1)]:
    x,yy=TL._images(*shp,5); print('gmsd strips',shp,TL.run_gmsd(lib,x,yy)[0], flush=True)
for shp in [(33,33,1,1),(64,43,2,1),(75,97,1,1)]:
    x,yy=TL._images(*shp,6); print('ssim4',shp,TL.run_ssim(lib,x,yy)[0], TL.run_ssim(lib,x,yy,taps=[0.2]*5)[0], TL.run_ssim(lib,x,yy,taps=[0.25]*4)[0], flush=True)
for (shp,L1,L2) in [((24,20,1,1),5,5),((45,70,1,1),11,3),((40,36,1,1),11,11),((12,9,1,1),1,4)]:
    x,yy=TL._images(*shp,7); print('ssim window',shp,L1,L2,TL.run_ssim_window(lib,x,yy,TL._window(L1,L2,L1+L2))[0], flush=True)
for (M,N,Pn,pads) in [(9,7,3,(5,5,5,5)),(6,11,2,(2,1,0,3)),(4,4,1,(4,4,4,4))]:
    src=np.asfortranarray(rng_pad:=np.random.default_rng(1).standard_normal((M,N,Pn,1)).astype(np.float32))
    out=np.asfortranarray(np.zeros((M+pads[0]+pads[1],N+pads[2]+pads[3],Pn,1),np.float32))
    lib.pad_symmetric(M,N,Pn,pads,0,src.ctypes.data,out.ctypes.data)
    back=np.asfortranarray(np.zeros((M,N,Pn,1),np.float32))
    lib.pad_symmetric_adjoint(M,N,Pn,pads,0,out.ctypes.data,back.ctypes.data)
    print('pad_symmetric',M,N,pads,float(np.abs(back).sum())>0, flush=True)
rng=np.random.default_rng(0)
for (M,N,C,B) in [(40,33,3,2),(7,5,5,1)]:
    img=rng.integers(0,256,size=(B,M,N,C),dtype=np.uint8); dst=np.asfortranarray(np.zeros((M,N,C,B),np.float32))
    lib.batch_from_n0f8(M,N,C,B,0,img.ctypes.data,1,C*N,C,M*N*C,dst.ctypes.data)
    assert np.array_equal(np.array(dst), img.transpose(1,2,3,0).astype(np.float32)/np.float32(255))
print('losses / batch ok')

# round 2: k_small (whole-call persistent kernel, >= 64 plane pairs), per-iteration parameters, isotropic fixed-order norm
for (M,P,B,k,K) in [(32,1,128,3,3),(64,1,128,0,2)]:
    y,h,_=make_case(M,M,P,B,k,k,70+M)
    print('small', M, check_forward(be,y,h,0.02,0.1,False,K), flush=True)
y,h,_=make_case(32,64,2,2,3,3,81)
xbar=torch.from_numpy(np.random.default_rng(2).standard_normal((32,64,2,2)))
for iso in (False, True):
    print('per-iteration', iso, check_backward(be,y,h,np.array([0.02,0.03,0.04],np.float32),np.array([0.1,0.2,0.15],np.float32),iso,3,xbar,flags=1,tol=1e-5,tol_scalar=2e-4), flush=True)
print('round-2 cases ok')
