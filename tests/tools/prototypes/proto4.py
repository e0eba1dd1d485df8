import sys; sys.path.insert(0,'/root/repo/tests/tools/prototypes')
from proto import *
import pickle
from oracle.condensed_model import _rows_apply, _rows_apply_T, _assemble_K
insts=pickle.load(open('/root/repo/gpurun_out/bad_T8.pkl','rb'))
def ipm_trace(c, frac_fn, max_iter=40, mu_tol=1e-13, verbose=True, guard=None):
    T = len(c.q)//2; n=2*T
    hi = np.stack([c.hi[:T], c.hi[T:2*T], np.append(c.hi[2*T:3*T-1], 1.0), c.hi[3*T-1:]])
    lo = np.stack([c.lo[:T], c.lo[T:2*T], np.append(c.lo[2*T:3*T-1], -1.0), c.lo[3*T-1:]])
    live = np.ones((4,T),bool); live[2,T-1]=False; nrow=2*live.sum()
    u=np.zeros(n); z=_rows_apply(T,u)
    sh=np.maximum(hi-z,1e-2); sl=np.maximum(z-lo,1e-2); lh=np.where(live,1.0,0.0); ll=lh.copy()
    gscale=1+np.abs(c.q).max()
    for it in range(1,max_iter+1):
        z=_rows_apply(T,u)
        rd=c.P@u+c.q+_rows_apply_T(T,np.where(live,lh-ll,0.0))
        rph=np.where(live,z+sh-hi,0.0); rpl=np.where(live,-z+sl+lo,0.0)
        mu=float((lh*sh+ll*sl)[live].sum())/nrow
        if mu<=mu_tol and max(np.abs(rph).max(),np.abs(rpl).max())<=1e-9 and np.abs(rd).max()<=1e-9*gscale: return u,it-1,True
        w=np.where(live,lh/sh+ll/sl,0.0)
        Lc=np.linalg.cholesky(_assemble_K(T,c.P,w))
        def newton(rch,rcl):
            th=np.where(live,(-rch+lh*rph)/sh,0.0); tl=np.where(live,(-rcl+ll*rpl)/sl,0.0)
            du=np.linalg.solve(Lc.T,np.linalg.solve(Lc,-rd-_rows_apply_T(T,th-tl)))
            dz=_rows_apply(T,du); dsh=-rph-dz; dsl=-rpl+dz
            dlh=np.where(live,(-rch-lh*dsh)/sh,0.0); dll=np.where(live,(-rcl-ll*dsl)/sl,0.0)
            return du,dsh,dsl,dlh,dll
        def raw(v,dv):
            m=live&(dv<0)
            return float((-v[m]/dv[m]).min()) if m.any() else np.inf
        du,dsh,dsl,dlh,dll=newton(lh*sh,ll*sl)
        aa=min(1.0,raw(sh,dsh),raw(sl,dsl),raw(lh,dlh),raw(ll,dll))
        mu_aff=float(((lh+aa*dlh)*(sh+aa*dsh)+(ll+aa*dll)*(sl+aa*dsl))[live].sum())/nrow
        sigma=(mu_aff/mu)**3
        du,dsh,dsl,dlh,dll=newton(lh*sh+dsh*dlh-sigma*mu, ll*sl+dsl*dll-sigma*mu)
        amax=min(raw(sh,dsh),raw(sl,dsl),raw(lh,dlh),raw(ll,dll))
        alpha=min(1.0,(frac_fn(mu,aa) if frac_fn.__code__.co_argcount==2 else frac_fn(mu))*amax)
        if guard:
            # backtrack until every pair keeps s*l >= guard * mu_new
            for _ in range(20):
                nsh,nsl,nlh,nll=sh+alpha*dsh,sl+alpha*dsl,lh+alpha*dlh,ll+alpha*dll
                mun=float((nlh*nsh+nll*nsl)[live].sum())/nrow
                if min((nlh*nsh)[live].min(),(nll*nsl)[live].min())>=guard*mun: break
                alpha*=0.8
        u=u+alpha*du; sh=sh+alpha*dsh; sl=sl+alpha*dsl; lh=lh+alpha*dlh; ll=ll+alpha*dll
        if verbose: print(it,"mu %.2e sigma %.2e aa %.3f alpha %.4f  min(s*l)/mu %.2e"%(mu,sigma,aa,alpha,min((lh*sh)[live].min(),(ll*sl)[live].min())/mu))
    return u,max_iter,False
if __name__=="__main__":
    all300=pickle.load(open('/tmp/inst_c2_300.pkl','rb'))
    rules={"0.99":lambda mu:0.99,
           "aa>=0.9:0.999":lambda mu,aa: 0.999 if aa>=0.9 else 0.99,
           "aa>=0.7:0.999":lambda mu,aa: 0.999 if aa>=0.7 else 0.99,
           "aa>=0.5:max(.99,1-mu)c.9999":lambda mu,aa: min(0.9999,max(0.99,1-mu)) if aa>=0.5 else 0.99,
           "1-(1-aa)*0.1 clipped":lambda mu,aa: min(0.9999,max(0.99, 1-0.1*(1-aa)**2)) }
    for name,rule in rules.items():
        bad=[ipm_trace(cq,rule,verbose=False)[1:] for p,r,cq in insts]
        res=[ipm_trace(cq,rule,verbose=False) for p,r,cq in all300]
        its=np.array([x[1] for x in res]); errs=[ctrl_err(cq,r,x[0]) for x,(p,r,cq) in zip(res,all300)]
        print(name,"| bad instances:",bad,"| c2-300: mean %.2f max %d fails %d errmax %.1e"%(its.mean(),its.max(),sum(not x[2] for x in res),max(errs)))
