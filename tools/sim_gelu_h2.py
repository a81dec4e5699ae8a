import numpy as np
from scipy.special import erf
h=np.float16
def r(x): return x.astype(np.float16)
def fma(a,b,c): return r(a.astype(np.float32)*b.astype(np.float32)+c.astype(np.float32))
def mul(a,b): return r(a.astype(np.float32)*b.astype(np.float32))
C0,C1,C2=0.7974584707823228,0.037050345100392894,-0.000358732361145582
x32=np.concatenate([np.linspace(-10,10,400001), np.random.randn(400000)*1.5]).astype(np.float32)
x=r(x32)
xc=np.clip(x,h(-8),h(8))
x2=mul(xc,xc)
p=fma(x2,np.full_like(x,C2),np.full_like(x,C1)); p=fma(x2,p,np.full_like(x,C0))
u=mul(xc,p)
t=r(np.tanh(u.astype(np.float64)))     # tanh.approx.f16x2: ~ f16-accurate
hp=fma(t,np.full_like(x,0.5),np.full_like(x,0.5))
g=mul(x,hp)
omt2=fma(-t,t,np.full_like(x,1))
du=fma(x2,np.full_like(x,5*C2),np.full_like(x,3*C1)); du=fma(x2,du,np.full_like(x,C0))
q=mul(mul(xc,np.full_like(x,0.5)),omt2)
d=fma(q,du,hp)
xe=x32.astype(np.float64)
Phi=0.5*(1+erf(xe/np.sqrt(2))); ge=xe*Phi; de=Phi+xe*np.exp(-xe*xe/2)/np.sqrt(2*np.pi)
def bf16(a):
    b=a.astype(np.float32).view(np.uint32); b=((b+0x7fff+((b>>16)&1))&0xffff0000).astype(np.uint32); return b.view(np.float32)
g_bf=bf16(g.astype(np.float32)); d_bf=bf16(d.astype(np.float32))
ref_g=bf16(ge.astype(np.float32)); ref_d=bf16(de.astype(np.float32))
for name,a,b,e in (("g",g_bf,ref_g,ge),("d",d_bf,ref_d,de)):
    err=np.abs(a-e); errref=np.abs(b-e)
    print(name,"max abs err f16-path", err.max(), "max abs err exact-then-bf16", errref.max(), "rms", np.sqrt((err**2).mean()), np.sqrt((errref**2).mean()),
          "max rel (|e|>0.05)", (err/np.abs(e))[np.abs(e)>0.05].max(), (errref/np.abs(e))[np.abs(e)>0.05].max())
