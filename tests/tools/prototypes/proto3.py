"""IPM variants on cached condensed instances: separate step lengths, step fraction, initial point."""
import sys; sys.path.insert(0,'/root/repo/tests/tools/prototypes')
from proto import *
import pickle
from oracle.condensed_model import _rows_apply, _rows_apply_T, _assemble_K

def ipm_var(c, sep_steps=False, frac=0.99, adaptive_frac=False, init='zero', max_iter=40, mu_tol=1e-13, gondzio=0):
    T = len(c.q)//2; n=2*T
    hi = np.stack([c.hi[:T], c.hi[T:2*T], np.append(c.hi[2*T:3*T-1], 1.0), c.hi[3*T-1:]])
    lo = np.stack([c.lo[:T], c.lo[T:2*T], np.append(c.lo[2*T:3*T-1], -1.0), c.lo[3*T-1:]])
    live = np.ones((4,T),bool); live[2,T-1]=False; nrow=2*live.sum()
    u=np.zeros(n)
    if init=='uncon':
        u=np.linalg.solve(c.P + 1e-3*np.eye(n)*np.abs(np.diag(c.P)).max(), -c.q)
        z=_rows_apply(T,u); u = u*0.0 if False else u
    z=_rows_apply(T,u)
    sh=np.maximum(hi-z,1e-2); sl=np.maximum(z-lo,1e-2)
    lh=np.where(live,1.0,0.0); ll=lh.copy()
    gscale=1+np.abs(c.q).max()
    for it in range(1,max_iter+1):
        z=_rows_apply(T,u)
        rd=c.P@u+c.q+_rows_apply_T(T,np.where(live,lh-ll,0.0))
        rph=np.where(live,z+sh-hi,0.0); rpl=np.where(live,-z+sl+lo,0.0)
        mu=float((lh*sh+ll*sl)[live].sum())/nrow
        if mu<=mu_tol and max(np.abs(rph).max(),np.abs(rpl).max())<=1e-9 and np.abs(rd).max()<=1e-9*gscale: return u,it-1,True
        w=np.where(live,lh/sh+ll/sl,0.0)
        K=_assemble_K(T,c.P,w)
        try: Lc=np.linalg.cholesky(K)
        except np.linalg.LinAlgError: return u,it,False
        def newton(rch,rcl):
            th=np.where(live,(-rch+lh*rph)/sh,0.0); tl=np.where(live,(-rcl+ll*rpl)/sl,0.0)
            du=np.linalg.solve(Lc.T,np.linalg.solve(Lc,-rd-_rows_apply_T(T,th-tl)))
            dz=_rows_apply(T,du); dsh=-rph-dz; dsl=-rpl+dz
            dlh=np.where(live,(-rch-lh*dsh)/sh,0.0); dll=np.where(live,(-rcl-ll*dsl)/sl,0.0)
            return du,dsh,dsl,dlh,dll
        def ms(v,dv):
            m=live&(dv<0)
            return min(1.0,float((-v[m]/dv[m]).min())) if m.any() else 1.0
        du,dsh,dsl,dlh,dll=newton(lh*sh,ll*sl)
        ap=min(ms(sh,dsh),ms(sl,dsl)); ad=min(ms(lh,dlh),ms(ll,dll))
        if not sep_steps: ap=ad=min(ap,ad)
        mu_aff=float(((lh+ad*dlh)*(sh+ap*dsh)+(ll+ad*dll)*(sl+ap*dsl))[live].sum())/nrow
        sigma=(mu_aff/mu)**3
        du,dsh,dsl,dlh,dll=newton(lh*sh+dsh*dlh-sigma*mu, ll*sl+dsl*dll-sigma*mu)
        f=frac
        if adaptive_frac: f=max(frac,1-mu) if mu<1 else frac
        ap=min(1.0,f*min(ms_raw(sh,dsh,live),ms_raw(sl,dsl,live))); ad=min(1.0,f*min(ms_raw(lh,dlh,live),ms_raw(ll,dll,live)))
        if not sep_steps: ap=ad=min(ap,ad)
        u=u+ap*du; sh=sh+ap*dsh; sl=sl+ap*dsl; lh=lh+ad*dlh; ll=ll+ad*dll
    return u,max_iter,False
def ms_raw(v,dv,live):
    m=live&(dv<0)
    return float((-v[m]/dv[m]).min()) if m.any() else np.inf

if __name__=="__main__":
    inst=pickle.load(open('/tmp/inst_c2_300.pkl','rb'))
    for kw in [dict(adaptive_frac=True,sep_steps=True), dict(adaptive_frac=True,frac=0.999), dict(adaptive_frac=True,frac=0.999,sep_steps=True), dict(adaptive_frac=True,frac=0.9999)]:
        its=[];errs=[];ok=[]
        for p,r,cq in inst:
            u,it,o=ipm_var(cq,**kw); its.append(it); errs.append(ctrl_err(cq,r,u)); ok.append(o)
        its=np.array(its); errs=np.array(errs)
        print(kw,"iters mean %.2f p99 %d max %d | err max %.2e | fails %d"%(its.mean(),np.percentile(its,99),its.max(),errs.max(),len(ok)-sum(ok)))
