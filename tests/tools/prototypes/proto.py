"""Prototype solvers on the condensed QP (numpy) - exploring iteration counts. Not product code."""
import sys, time; sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/av-simulation-at-intersections_b200')
import numpy as np
from oracle import mpc_oracle as O, condensed_model as CM
from junction_mpc import synth
from junction_mpc.config import PARAM_INDEX as PI

def params_from_vec(v, T):
    return O.Params(T=T, dt=v[PI['dt']], dl=v[PI['dl']], L=v[PI['L']], speed=v[PI['speed']], w_perp=v[PI['w_perp']], w_para=v[PI['w_para']],
        R=(v[PI['R_a']], v[PI['R_d']]), Rd=(v[PI['Rd_a']], v[PI['Rd_d']]), Q_v_yaw=(v[PI['Q_v']], v[PI['Q_yaw']]),
        Qf=(v[PI['Qf_x']],v[PI['Qf_y']],v[PI['Qf_v']],v[PI['Qf_yaw']]), R_end=(v[PI['Rend_a']],v[PI['Rend_d']]),
        max_dsteer=v[PI['max_dsteer']], max_accel=v[PI['max_accel']], max_decel=v[PI['max_decel']], max_steer=v[PI['max_steer']],
        sim_max_speed=v[PI['sim_max_speed']], min_speed=v[PI['min_speed']], v_ref_min=v[PI['v_ref_min']])

def instances(w, count):
    from junction_mpc.config import MPCConfig
    T=w['T']; c=w['courses'][0]
    base = MPCConfig.default().with_T(T).param_vector(dl=w['dl'])
    out=[]
    for k in range(count):
        pv = base if w['params'] is None else w['params'][:,k]
        p = params_from_vec(pv, T)
        n=w['course_len'][k]; x0=w['state'][:,k]
        r = O.mpc_step(p, x0, w['oa'][:,k], w['od'][:,k], c[:n,0], c[:n,1], c[:n,2], int(w['target_ind'][k]))
        if r.status!=0: print("oracle status", r.status, r.qp.kkt if r.qp else None); continue
        cq = CM.condense(p, r.xref, r.xbar, x0, r.reaches_end)
        out.append((p, r, cq))
    return out

def ctrl_err(cq, r, u):
    X = CM.states_from_controls(cq,u); T=len(r.oa)
    def e(a,b): return np.max(np.abs(a-b) - 1e-3*np.abs(b))
    return max(e(u[:T],r.oa), e(u[T:],r.od), e(X[0],r.ox), e(X[1],r.oy), e(X[2],r.ov), e(X[3],r.oyaw))

def ipm(cq, max_iter=40, tol_mu=float(__import__("os").environ.get("TOLMU","1e-11")), track=None):
    """Mehrotra PC on condensed QP with two-sided rows: lo <= A u <= hi."""
    P,q,A,lo,hi = cq.P,cq.q,cq.A,cq.lo,cq.hi
    n=len(q); m=len(lo)
    G=np.vstack([A,-A]); h=np.concatenate([hi,-lo])
    u=np.zeros(n)
    # start: u = clip(0) interior
    s=h-G@u; s=np.maximum(s,1e-2)   # infeasible start ok
    lam=np.ones(2*m)
    hist=[]
    for it in range(1,max_iter+1):
        rd=P@u+q+G.T@lam; rp=G@u+s-h; mu=lam@s/(2*m)
        W=lam/s
        K=P+G.T@(W[:,None]*G)
        L=np.linalg.cholesky(K)
        def solve(rhs): return np.linalg.solve(L.T, np.linalg.solve(L,rhs))
        def newton(rc):
            t=(-rc+lam*rp)/s
            du=solve(-rd-G.T@t); ds=-rp-G@du; dl=(-rc-lam*ds)/s
            return du,ds,dl
        def mstep(v,dv):
            neg=dv<0
            return min(1.0, (-v[neg]/dv[neg]).min()) if neg.any() else 1.0
        du,ds,dl=newton(lam*s)
        aa=min(mstep(s,ds),mstep(lam,dl))
        mua=(lam+aa*dl)@(s+aa*ds)/(2*m)
        sig=(mua/mu)**3
        du,ds,dl=newton(lam*s+ds*dl-sig*mu)
        a=min(1.0, 0.99*min(mstep(s,ds),mstep(lam,dl)))
        u=u+a*du; s=s+a*ds; lam=lam+a*dl
        if track is not None: hist.append(track(u))
        rd=P@u+q+G.T@lam; rp=G@u+s-h; mu=lam@s/(2*m)
        if mu<tol_mu and np.abs(rp).max()<1e-9 and np.abs(rd).max()<1e-7*(1+np.abs(q).max()): break
    return u,it,hist

if __name__=="__main__":
    which=sys.argv[1] if len(sys.argv)>1 else 'c2'
    cnt=int(sys.argv[2]) if len(sys.argv)>2 else 100
    if which=='c2': w=synth.make_workload(2,B=cnt)
    else: w=synth.make_sweep(int(which[1:]), states_per_point=1, max_points=cnt)
    inst=instances(w,cnt)
    its=[]; errs=[]
    for p,r,cq in inst:
        u,it,hist=ipm(cq, track=lambda u: ctrl_err(cq,r,u))
        its.append(it); errs.append(ctrl_err(cq,r,u))
        first_ok = next((i+1 for i,e in enumerate(hist) if e<1e-4 and all(x<1e-4 for x in hist[i:])), None)
        its[-1]=(it,first_ok)
    a=np.array([i[0] for i in its]); f=np.array([i[1] if i[1] else 99 for i in its])
    print("IPM iters: mean %.1f max %d ; first iteration within parity: mean %.1f max %d ; final err max %.2e  (#>1e-4: %d)"%(a.mean(),a.max(),f.mean(),f.max(),max(errs),sum(e>1e-4 for e in errs)))
