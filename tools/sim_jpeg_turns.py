"""Simulation behind profiles/r02_experiments.md: warp turns of the JPEG AC loop per CTA (bench content, q = 90) as mapped now, with the CTA's blocks sorted by non-zero count, and ideal."""
import numpy as np
W,H=7680,1024   # a band of the frame is enough
yy,xx=np.mgrid[0:H,0:W]
rng=np.random.default_rng(0)
rgb=np.stack([xx*255//(7680-1), yy*255//(4320-1), (xx+yy)%256],axis=2).astype(np.int32)+rng.integers(-6,7,(H,W,3))
rgb=rgb.clip(0,255).astype(np.float64)
# BT.601 limited as UltraGrid's RGB->UYVY roughly (enough for statistics)
r,g,b=rgb[...,0],rgb[...,1],rgb[...,2]
Y=np.round(0.257*r+0.504*g+0.098*b+16); Cb=np.round(-0.148*r-0.291*g+0.439*b+128); Cr=np.round(0.439*r-0.368*g-0.071*b+128)
Cb=np.floor((Cb[:,0::2]+Cb[:,1::2])/2); Cr=np.floor((Cr[:,0::2]+Cr[:,1::2])/2)
ql=np.array([16,11,10,16,24,40,51,61,12,12,14,19,26,58,60,55,14,13,16,24,40,57,69,56,14,17,22,29,51,87,80,62,18,22,37,56,68,109,103,77,24,35,55,64,81,104,113,92,49,64,78,87,103,121,120,101,72,92,95,98,112,100,103,99]).reshape(8,8)
qc=np.array([17,18,24,47,99,99,99,99,18,21,26,66,99,99,99,99,24,26,56,99,99,99,99,99,47,66,99,99,99,99,99,99]+[99]*32).reshape(8,8)
def scale(q,Q=90):
    s=200-2*Q
    return np.clip((q*s+50)//100,1,255)
ql,qc=scale(ql),scale(qc)
k=np.arange(8); C=np.cos((2*k[None,:]+1)*k[:,None]*np.pi/16)*np.sqrt(2/8); C[0]/=np.sqrt(2)
def blocks(P,q):
    h,w=P.shape
    B=(P-128).reshape(h//8,8,w//8,8).transpose(0,2,1,3)
    D=np.einsum('ij,abjk,lk->abil',C,B,C)
    return np.round(D/q).astype(int)
zz=np.array([0,1,8,16,9,2,3,10,17,24,32,25,18,11,4,5,12,19,26,33,40,48,41,34,27,20,13,6,7,14,21,28,35,42,49,56,57,50,43,36,29,22,15,23,30,37,44,51,58,59,52,45,38,31,39,46,53,60,61,54,47,55,62,63])
def nnz(Q):
    f=Q.reshape(Q.shape[0],Q.shape[1],64)[...,zz]
    lo=(f[...,1:32]!=0).sum(-1); hi=(f[...,32:]!=0).sum(-1)
    return lo,hi
ylo,yhi=nnz(blocks(Y,ql)); blo,bhi=nnz(blocks(Cb,qc)); rlo,rhi=nnz(blocks(Cr,qc))
print("avg nnz Y %.2f Cb %.2f Cr %.2f"%((ylo+yhi).mean(),(blo+bhi).mean(),(rlo+rhi).mean()))
# CTA = 32 consecutive MCUs in a block row; MCU = Y0 Y1 Cb Cr
bh,bwY=ylo.shape
cur=[];srt=[];avg=[]
for by in range(bh):
    for m0 in range(0,bwY//2,32):
        lo=np.stack([ylo[by,2*m0:2*m0+64:2],ylo[by,2*m0+1:2*m0+64:2],blo[by,m0:m0+32],rlo[by,m0:m0+32]])  # [4 warps][32 lanes]
        hi=np.stack([yhi[by,2*m0:2*m0+64:2],yhi[by,2*m0+1:2*m0+64:2],bhi[by,m0:m0+32],rhi[by,m0:m0+32]])
        cur.append((lo.max(1)+hi.max(1)).sum())
        tot=(lo+hi).reshape(-1); o=np.argsort(tot)
        l2=lo.reshape(-1)[o].reshape(4,32); h2=hi.reshape(-1)[o].reshape(4,32)
        srt.append((l2.max(1)+h2.max(1)).sum())
        avg.append(tot.sum()/32)
print("per CTA: warp-iterations current %.1f  sorted %.1f  ideal(balanced) %.1f"%(np.mean(cur),np.mean(srt),np.mean(avg)))
cur1=[];srt1=[]
for by in range(bh):
    for m0 in range(0,bwY//2,32):
        lo=np.stack([ylo[by,2*m0:2*m0+64:2],ylo[by,2*m0+1:2*m0+64:2],blo[by,m0:m0+32],rlo[by,m0:m0+32]])
        hi=np.stack([yhi[by,2*m0:2*m0+64:2],yhi[by,2*m0+1:2*m0+64:2],bhi[by,m0:m0+32],rhi[by,m0:m0+32]])
        tot=(lo+hi)
        cur1.append(tot.max(1).sum())
        srt1.append(np.sort(tot.reshape(-1)).reshape(4,32).max(1).sum())
print("single 64-bit loop: current %.1f sorted %.1f"%(np.mean(cur1),np.mean(srt1)))
print("hi-half nonzero fraction of blocks: Y %.3f"%((yhi>0).mean()))
