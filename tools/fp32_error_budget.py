#!/usr/bin/env python3
"""CPU emulation of the fp32 BS_MATH_FAST formula (numpy float32, fma emulated in float64) with exact or perturbed
elementary operations: where does the distance to the reference fp32 output come from?  """
import sys, numpy as np
sys.path.insert(0,'/root/repo/tests'); sys.path.insert(0,'/root/repo')
import oracle_lib
from gpu_util import inputgen_like
f32=np.float32
n=4_000_000
s,k,r,v,t,o=inputgen_like(n,seed=78)
ref=oracle_lib.price_map(s,k,r,v,t,o,4).astype(np.float64)
ref64=oracle_lib.price_map(s.astype(np.float64),k.astype(np.float64),r.astype(np.float64),v.astype(np.float64),t.astype(np.float64),o,8)
print("ref32 vs ref64 max", np.abs(ref-ref64).max())
def fma(a,b,c): return (a.astype(np.float64)*b.astype(np.float64)+c.astype(np.float64)).astype(f32)
def fast(ex2_err=0.0, lg2_err=0.0, rcp_err=0.0, seed=1, exact_fv=False, exact_cndf=False, exact_rcp=False):
    rng=np.random.RandomState(seed)
    def ex2(x, exact=False):
        y=np.exp2(x.astype(np.float64))
        if not exact and ex2_err: y=y*(1+rng.uniform(-ex2_err,ex2_err,x.shape))
        return y.astype(f32)
    def lg2(x):
        y=np.log2(x.astype(np.float64))
        if lg2_err: y=y+rng.uniform(-lg2_err,lg2_err,x.shape)
        return y.astype(f32)
    def rcp(x, exact=False):
        y=1.0/x.astype(np.float64)
        if not exact and rcp_err: y=y*(1+rng.uniform(-rcp_err,rcp_err,x.shape))
        return y.astype(f32)
    LN2=f32(0.69314718055994530942); NL2E=f32(-1.44269504088896340736); NH=f32(-0.72134752044448170368)
    sq=np.sqrt(t.astype(np.float64)).astype(f32)
    den=(v*sq).astype(f32); rden=rcp(den, exact_rcp)
    lg=(lg2(s)-lg2(k)).astype(f32)
    drift=fma((f32(0.5)*v).astype(f32),v,r)
    num=fma(lg,np.full(n,LN2,f32),(drift*t).astype(f32))
    d1=(num*rden).astype(f32); d2=(d1-den).astype(f32)
    fv=(k*ex2(((r*NL2E).astype(f32)*t).astype(f32), exact_fv)).astype(f32)
    A=[f32(c*0.39894228040143270286) for c in (0.319381530,-0.356563782,1.781477937,-1.821255978,1.330274429)]
    def tail(d):
        kk=rcp(fma(np.abs(d),np.full(n,f32(0.2316419)),np.full(n,f32(1.0))), exact_rcp)
        e=ex2(((d*NH).astype(f32)*d).astype(f32), exact_cndf)
        p=fma(kk,np.full(n,A[4]),np.full(n,A[3])); p=fma(kk,p,np.full(n,A[2])); p=fma(kk,p,np.full(n,A[1])); p=fma(kk,p,np.full(n,A[0]))
        return ((p*kk).astype(f32)*e).astype(f32)
    w1=tail(d1); w2=tail(d2)
    put=o!=0
    x1=np.where((d1<0)!=put, w1, (f32(1)-w1).astype(f32)); x2=np.where((d2<0)!=put, w2,(f32(1)-w2).astype(f32))
    c=fma(s,x1,-(fv*x2).astype(f32))
    return np.where(put,-c,c).astype(np.float64)
E=2.0**-22
for name,kw in [("exact ops",{}),("ex2 err",dict(ex2_err=E)),("lg2 err",dict(lg2_err=E)),("rcp err 2^-23",dict(rcp_err=2.0**-23)),
                ("all",dict(ex2_err=E,lg2_err=E,rcp_err=2.0**-23)),("all, exact fv",dict(ex2_err=E,lg2_err=E,rcp_err=2.0**-23,exact_fv=True)),
                ("all, exact fv+cndf exp",dict(ex2_err=E,lg2_err=E,rcp_err=2.0**-23,exact_fv=True,exact_cndf=True)),
                ("all, exact exps+rcp",dict(ex2_err=E,lg2_err=E,rcp_err=2.0**-23,exact_fv=True,exact_cndf=True,exact_rcp=True))]:
    p=fast(**kw)
    print("%-28s vs ref32 max %.3e   vs ref64 max %.3e"%(name, np.abs(p-ref).max(), np.abs(p-ref64).max()))
