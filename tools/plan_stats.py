import sys, os, ctypes as C
sys.path.insert(0,'/root/repo/tests'); sys.path.insert(0,'/root/repo')
import numpy as np
from qcsim_b200 import circuits, _lib
import test_planner as tp
hl = C.CDLL(tp.SO); hl.hl_plan.restype=C.c_int; hl.hl_rounds.restype=C.c_int
def analyze(circ, n, K=12, L=4, maxvar=2):
    N=len(circ); arr=tp.pack(circ)
    fused=(C.c_int*N)(); tile=(C.c_ulonglong*N)(); nops=(C.c_int*N)(); order=(C.c_int*N)(); ops=(tp.HlOp*N)()
    ns=hl.hl_plan(arr,N,n,K,L,fused,tile,nops,order,ops)
    pos=0; tot_rounds=0; passes=0; singles=0; launches=0
    for s in range(ns):
        members=list(order[pos:pos+nops[s]]); pos+=nops[s]
        if not fused[s]: singles+=1; continue
        sub=[circ[i] for i in members]; M=len(sub); a2=tp.pack(sub)
        R=M+4
        rb=(C.c_int*(3*R))(); nv=(C.c_int*R)(); vq=(C.c_int*(3*R))(); no=(C.c_int*R)(); od=(C.c_int*M)(); ib=(C.c_int*(9*R))()
        nr=hl.hl_rounds(a2,M,tile[s],maxvar,rb,nv,vq,no,od,ib)
        mats=sum(1<<nv[r] for r in range(nr))
        tot_rounds+=nr; passes+=1; launches+=max(-(-nr//7), -(-mats//28))
    return passes, launches, tot_rounds, singles
n=30
for layers in (1,10):
    circ=circuits.random_circuit(n, layers)
    for K,L in ((12,4),(12,3),(12,2)):
        for mv in (2,3):
            p,l,r,s=analyze(circ,n,K,L,mv)
            print(f"layers={layers} K={K} L={L} maxvar={mv}: passes={p} launches={l} rounds={r} singles={s}  per layer: {p/layers:.2f} passes, {r/layers:.1f} rounds")
