import sys; sys.path.insert(0,'/root/repo/tests/tools/prototypes')
from proto import *

def ruiz(P,A,iters=10):
    n=P.shape[0]; m=A.shape[0]
    D=np.ones(n); E=np.ones(m)
    for _ in range(iters):
        Ps=D[:,None]*P*D[None,:]; As=E[:,None]*A*D[None,:]
        cn=np.maximum(np.abs(Ps).max(0), np.abs(As).max(0)); cn[cn<1e-4]=1
        rn=np.abs(As).max(1); rn[rn<1e-4]=1
        D/=np.sqrt(cn); E/=np.sqrt(rn)
    Ps=D[:,None]*P*D[None,:]
    c=1/max(np.mean(np.abs(Ps).max(0)),1e-4)
    return D,E,c

def active_set(cq,u,y,tol=1e-6):
    z=cq.A@u
    low=(z-cq.lo < -y); upp=(cq.hi-z < y)
    return low,upp

def eqp(cq,low,upp):
    """exact solve with active set; returns u, multipliers, and violation info"""
    A=cq.A; n=len(cq.q)
    act=np.nonzero(low|upp)[0]
    b=np.where(low,cq.lo,cq.hi)[act]
    Aa=A[act]
    K=np.block([[cq.P,Aa.T],[Aa,-1e-12*np.eye(len(act))]])
    rhs=np.concatenate([-cq.q,b])
    sol=np.linalg.solve(K,rhs); sol+=np.linalg.solve(K,rhs-K@sol)
    u=sol[:n]; lam=np.zeros(A.shape[0]); lam[act]=sol[n:]
    return u,lam

def pdas(cq,low,upp,rounds=10):
    for r in range(rounds):
        u,lam=eqp(cq,low,upp)
        z=cq.A@u
        viol_lo=(z<cq.lo-1e-10)&~low; viol_hi=(z>cq.hi+1e-10)&~upp
        bad_lo=low&(lam>1e-10); bad_hi=upp&(lam<-1e-10)   # lower-active needs lam<=0 ; upper needs lam>=0
        if not (viol_lo.any() or viol_hi.any() or bad_lo.any() or bad_hi.any()): return u,r+1,True
        low=(low|viol_lo)&~bad_lo; upp=(upp|viol_hi)&~bad_hi
    return u,rounds,False

def admm(cq, r, max_iter=2000, rho0=0.1, sigma=1e-6, alpha=1.6, check=10, adapt=True, polish_every=25):
    P,q,A,lo,hi=cq.P,cq.q,cq.A,cq.lo,cq.hi
    n=len(q); m=len(lo)
    D,E,c=ruiz(P,A)
    Ps=c*D[:,None]*P*D[None,:]; qs=c*D*q; As=E[:,None]*A*D[None,:]; los=E*lo; his=E*hi
    rho=rho0
    x=np.zeros(n); z=np.zeros(m); y=np.zeros(m)
    def factor(rho): return np.linalg.cholesky(Ps+sigma*np.eye(n)+rho*As.T@As)
    L=factor(rho); nfac=1
    first_pol=None
    for k in range(1,max_iter+1):
        rhs=sigma*x-qs+As.T@(rho*z-y)
        xt=np.linalg.solve(L.T,np.linalg.solve(L,rhs))
        zt=As@xt
        xn=alpha*xt+(1-alpha)*x
        zh=alpha*zt+(1-alpha)*z
        zn=np.clip(zh+y/rho,los,his)
        y=y+rho*(zh-zn)
        x=xn; z=zn
        if k%check==0:
            Ax=As@x
            rp=np.abs((Ax-z)/E).max(); rd=np.abs((Ps@x+qs+As.T@y)/D).max()/c
            np_=max(np.abs(Ax/E).max(),np.abs(z/E).max()); nd=max(np.abs(Ps@x/D).max(),np.abs(As.T@y/D).max(),np.abs(qs/D).max())/c
            if k%polish_every==0:
                u=D*x; yy=E*y/c
                low,upp=active_set(cq,u,yy)
                up,rounds,ok=pdas(cq,low,upp,rounds=3)
                if ok and ctrl_err(cq,r,up)<1e-4:
                    return k,nfac,rounds,rp/(1e-12+np_),rd/(1e-12+nd)
            if adapt:
                ratio=np.sqrt((rp/(np_+1e-12))/(rd/(nd+1e-12)+1e-12))
                if ratio>5 or ratio<0.2:
                    rho=min(max(rho*ratio,1e-6),1e6); L=factor(rho); nfac+=1
    return max_iter,nfac,0,0,0

if __name__=="__main__":
    which=sys.argv[1]; cnt=int(sys.argv[2])
    if which=='c2': w=synth.make_workload(2,B=cnt)
    else: w=synth.make_sweep(int(which[1:]), states_per_point=1, max_points=cnt)
    inst=instances(w,cnt)
    res=[admm(cq,r) for p,r,cq in inst]
    k=np.array([x[0] for x in res]); f=np.array([x[1] for x in res]); rr=np.array([x[2] for x in res])
    print("ADMM iters until polish(<=3 PDAS rounds) gives parity: mean %.0f median %.0f p90 %.0f max %d ; refactors mean %.1f; pdas rounds mean %.2f; fail %d"%(k.mean(),np.median(k),np.percentile(k,90),k.max(),f.mean(),rr.mean(),(k>=2000).sum()))
    # cold PDAS
    cold=[pdas(cq,np.zeros(len(cq.lo),bool),np.zeros(len(cq.lo),bool),rounds=30) for p,r,cq in inst]
    okc=[c[2] and ctrl_err(cq,r,c[0])<1e-4 for c,(p,r,cq) in zip(cold,inst)]
    print("cold PDAS: success %d/%d rounds mean %.1f max %d"%(sum(okc),len(okc),np.mean([c[1] for c in cold]),max(c[1] for c in cold)))
