import sys; sys.path.insert(0,'/root/repo/tests/tools/prototypes')
from proto import *
from multiprocessing import Pool
def work(args):
    which,cnt,lo,hi=args
    if which=='c2': w=synth.make_workload(2,B=cnt)
    else: w=synth.make_sweep(int(which[1:]), states_per_point=1, max_points=cnt)
    for k in ['state','oa','od']: w[k]=w[k][:,lo:hi]
    for k in ['course_len','target_ind']: w[k]=w[k][lo:hi]
    if w['params'] is not None: w['params']=w['params'][:,lo:hi]
    inst=instances(w,hi-lo)
    out=[]
    for p,r,cq in inst:
        try:
            u,it,_=ipm(cq)
            out.append((it,ctrl_err(cq,r,u)))
        except Exception as e:
            out.append((-1,1.0))
    return out
if __name__=="__main__":
    which=sys.argv[1]; cnt=int(sys.argv[2])
    chunks=[(which,cnt,i,min(cnt,i+cnt//8+1)) for i in range(0,cnt,cnt//8+1)]
    with Pool(8) as pool: res=sum(pool.map(work,chunks),[])
    it=np.array([r[0] for r in res]); e=np.array([r[1] for r in res])
    print(which,"n",len(res),"iters mean %.1f p99 %d max %d; err max %.2e p99 %.2e; >1e-4: %d; fails %d"%(it.mean(),np.percentile(it,99),it.max(),e.max(),np.percentile(e,99),(e>1e-4).sum(),(it<0).sum()))
