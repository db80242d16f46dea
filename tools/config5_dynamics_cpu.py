#!/usr/bin/env python
"""Plain torch on the CPU, no part of this repo: the config-5 `random` row (20 000 x 20 000, d = 2, p = 0.1, batch
65536, lr 1e-2, 10 epochs) with the reference's step (gather, sigmoid, BCE, dense Adam with coupled L2).
With weight_decay = 1e-5 the tables collapse to zero within two epochs (loss 0.7077 -> 0.6931 = ln 2, the numbers
profiles/r02_config5.json shows for the GPU path); with weight_decay = 0 the same run reaches the ground-truth
accuracy ceiling (0.61).  So the flat rows of that record are the optimisation problem at that batch size, not the
kernels.  Usage: python tools/config5_dynamics_cpu.py [wd]   (about a minute per run)"""
import sys, math, torch
torch.manual_seed(0)
n=m=20000; d=2; p=0.1
A,_=torch.linalg.qr(torch.randn(n,d)); Bm,_=torch.linalg.qr(torch.randn(m,d)); scale=math.sqrt(n*m)/(2*math.sqrt(d))
N=int(n*m*p/2)
u=torch.randint(0,n,(N,)); i=torch.randint(0,m,(N,)); j=torch.randint(0,m,(N,))
dx=scale*((A[u]*Bm[i]).sum(1)-(A[u]*Bm[j]).sum(1))
z=(torch.rand(N)<torch.sigmoid(dx)).float()
print("gt acc", ((dx>0).float()==z).float().mean().item(), "std dX", dx.std().item())
ntr=int(0.8*N)
def run(lr,B,epochs,wd=1e-5):
    torch.manual_seed(1)
    U=(torch.randn(n,d)/d**0.5).requires_grad_(); V=(torch.randn(m,d)/d**0.5).requires_grad_()
    opt=torch.optim.Adam([U,V],lr=lr,weight_decay=wd)
    for e in range(epochs):
        perm=torch.randperm(ntr); tot=0; nb=0
        for s0 in range(0,ntr,B):
            k=perm[s0:s0+B]
            x=(U[u[k]]*(V[i[k]]-V[j[k]])).sum(1)
            loss=torch.nn.functional.binary_cross_entropy(torch.sigmoid(x), z[k])
            opt.zero_grad(); loss.backward(); opt.step(); tot+=loss.item(); nb+=1
        with torch.no_grad():
            k=torch.arange(ntr,N); x=(U[u[k]]*(V[i[k]]-V[j[k]])).sum(1)
            acc=((x>0).float()==z[k]).float().mean().item()
        print(f"lr {lr} B {B} epoch {e}: loss {tot/nb:.4f} test acc {acc:.4f} |U| {U.abs().mean().item():.4f}", flush=True)
run(0.01, 65536, 6, wd=float(sys.argv[1]) if len(sys.argv) > 1 else 1e-5)
