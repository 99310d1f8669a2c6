"""Evidence for the jerk tolerance in tests/helpers.py: CPU emulation (numpy float32) of the
library's FP32 pair arithmetic on double-single positions, on the golden set
tests/golden/ph4_plummer1k_eps1e-4.npz, against the reference ph4 FP64 forces stored there.

Part 1 varies the rsqrt error and a hypothetical FP64 near-field cut-off; part 2 promotes one
stage at a time to FP64.  Finding: with EVERYTHING after the FP32 rounding of dx = (xj.hi-xi.hi)
+(xj.lo-xi.lo) done in FP64 ("all"), max |d jerk|/|jerk| is still ~1e-6 and max |d acc|/|acc| ~1e-7
at N=1k: jerk is a sum of random-sign terms, so a few particles per thousand have |jerk| ~10x below
the size of the terms.  The raw MUFU.RSQ error (2^-22.9) costs another factor ~2.5, which is why
the library refines it with one Newton step by default.  Run: python tools/fp32_floor_emulation.py
"""
import os
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
g=dict(np.load(os.path.join(ROOT, 'tests/golden/ph4_plummer1k_eps1e-4.npz')))
x=g['pos']; v=g['vel']; m=g['mass']; eps2=np.float32(g['eps2'])
f=np.float32
xh=x.astype(f); xl=(x-xh.astype(np.float64)).astype(f)
def run(rsq_err=0.0, vds=False, rcut2=0.0, seed=0):
    rnd=np.random.RandomState(seed)
    n=len(m)
    vf=v.astype(f); vl=(v-vf.astype(np.float64)).astype(f)
    dx=[(xh[None,:,k]-xh[:,None,k])+(xl[None,:,k]-xl[:,None,k]) for k in range(3)]
    if vds:
        dv=[((vf[None,:,k]-vf[:,None,k])-vl[:,None,k]) for k in range(3)]
    else:
        dv=[(vf[None,:,k]-vf[:,None,k]) for k in range(3)]
    r2=dx[0]*dx[0]; r2=(dx[1].astype(np.float64)*dx[1]+r2).astype(f); r2=(dx[2].astype(np.float64)*dx[2]+r2).astype(f)
    xv=dx[0]*dv[0]; xv=(dx[1].astype(np.float64)*dv[1]+xv).astype(f); xv=(dx[2].astype(np.float64)*dv[2]+xv).astype(f)
    ok=(r2>2.2e-16)&(~np.eye(n,dtype=bool))
    far=ok&(r2>rcut2)
    r2e=(r2+eps2).astype(f)
    rinv=(1/np.sqrt(r2e.astype(np.float64)))
    rinv=(rinv*(1+rsq_err*rnd.uniform(-1,1,rinv.shape))).astype(f)
    rinv=np.where(far,rinv,f(0))
    rinv2=rinv*rinv; mrinv=m.astype(f)[None,:]*rinv; mr3=mrinv*rinv2
    a3=f(-3)*(xv*rinv2)
    acc=np.zeros((n,3)); jerk=np.zeros((n,3))
    for k in range(3):
        acc[:,k]=(mr3*dx[k]).astype(np.float64).sum(axis=1)
        t=(a3.astype(np.float64)*dx[k]+dv[k]).astype(f)
        jerk[:,k]=(mr3*t).astype(np.float64).sum(axis=1)
    pot=-(mrinv.astype(np.float64)).sum(axis=1)
    # near pairs in FP64
    near=ok&(~far)
    X=x[None,:,:]-x[:,None,:]; V=v[None,:,:]-v[:,None,:]
    R2=(X*X).sum(2); XV=(X*V).sum(2)
    r2i=1/(R2+float(eps2)+2.2e-16); ri=np.sqrt(r2i); mri=m[None,:]*ri; mr3i=mri*r2i; A3=-3*XV*r2i
    w=np.where(near,1.0,0.0)
    acc+=(w*mr3i)[:,:,None].__mul__(X).sum(1)
    jerk+=((w*mr3i)[:,:,None]*(V+A3[:,:,None]*X)).sum(1)
    pot-=(w*mri).sum(1)
    ea=np.linalg.norm(acc-g['acc'],axis=1)/np.linalg.norm(g['acc'],axis=1)
    ej=np.linalg.norm(jerk-g['jerk'],axis=1)/np.linalg.norm(g['jerk'],axis=1)
    ep=np.abs(pot-g['pot'])/np.abs(g['pot'])
    return ea.max(), np.percentile(ea,99), ej.max(), np.percentile(ej,99), ep.max(), near.sum()/n
for name,kw in [('ideal rsqrt',dict()),('rsqrt 1.3e-7',dict(rsq_err=1.3e-7)),('ideal+vds',dict(vds=True)),
                ('rcut 0.05 rsq',dict(rsq_err=1.3e-7,rcut2=0.05**2)),('rcut 0.1 rsq',dict(rsq_err=1.3e-7,rcut2=0.1**2)),
                ('rcut 0.1 rsq vds',dict(rsq_err=1.3e-7,rcut2=0.1**2,vds=True)),('rcut 0.2 rsq vds',dict(rsq_err=1.3e-7,rcut2=0.2**2,vds=True))]:
    print(name, ['%.2e'%t for t in run(**kw)])
print('---- component study (ideal rsqrt)')
def run2(vel64=False, xv64=False, t64=False, mr364=False, r264=False):
    n=len(m)
    F=np.float64
    vf=v if vel64 else v.astype(f)
    dx=[(xh[None,:,k]-xh[:,None,k])+(xl[None,:,k]-xl[:,None,k]) for k in range(3)]
    dv=[(vf[None,:,k]-vf[:,None,k]) for k in range(3)]
    if not vel64: dv=[d.astype(f) for d in dv]
    if r264:
        r2=sum(d.astype(F)*d for d in dx)
    else:
        r2=dx[0]*dx[0]; r2=(dx[1].astype(F)*dx[1]+r2).astype(f); r2=(dx[2].astype(F)*dx[2]+r2).astype(f)
    if xv64:
        xv=sum(dx[k].astype(F)*dv[k] for k in range(3))
    else:
        xv=(dx[0]*dv[0]).astype(f); xv=(dx[1].astype(F)*dv[1]+xv).astype(f); xv=(dx[2].astype(F)*dv[2]+xv).astype(f)
    ok=(r2>2.2e-16)&(~np.eye(n,dtype=bool))
    r2e=r2+(F(eps2) if r264 else eps2)
    if not r264: r2e=r2e.astype(f)
    rinv=1/np.sqrt(r2e.astype(F))
    if not mr364: rinv=rinv.astype(f)
    rinv=np.where(ok,rinv,0)
    if mr364:
        rinv2=rinv*rinv; mr3=m[None,:]*rinv*rinv2
    else:
        rinv2=(rinv*rinv).astype(f); mrinv=(m.astype(f)[None,:]*rinv).astype(f); mr3=(mrinv*rinv2).astype(f)
    if t64:
        a3=-3*(xv.astype(F)*rinv2)
    else:
        a3=(f(-3)*(xv.astype(f)*rinv2.astype(f)).astype(f)).astype(f)
    jerk=np.zeros((n,3)); acc=np.zeros((n,3))
    for k in range(3):
        if t64: t=a3*dx[k]+dv[k]
        else: t=(a3.astype(F)*dx[k]+dv[k]).astype(f)
        jerk[:,k]=(mr3.astype(F)*t).sum(axis=1)
        acc[:,k]=(mr3.astype(F)*dx[k]).sum(axis=1)
    ej=np.linalg.norm(jerk-g['jerk'],axis=1)/np.linalg.norm(g['jerk'],axis=1)
    ea=np.linalg.norm(acc-g['acc'],axis=1)/np.linalg.norm(g['acc'],axis=1)
    return '%.2e %.2e | acc %.2e'%(ej.max(), np.percentile(ej,99), ea.max())
print('base', run2())
print('vel64', run2(vel64=True))
print('xv64', run2(xv64=True))
print('t64', run2(t64=True))
print('mr364', run2(mr364=True))
print('r264', run2(r264=True))
print('vel64+xv64', run2(vel64=True,xv64=True))
print('vel64+xv64+t64', run2(vel64=True,xv64=True,t64=True))
print('all', run2(vel64=True,xv64=True,t64=True,mr364=True,r264=True))
