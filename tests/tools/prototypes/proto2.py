import sys; sys.path.insert(0,'/root/repo/tests/tools/prototypes')
from proto import *
import pickle, os
def get(which,cnt):
    fn=f'/tmp/inst_{which}_{cnt}.pkl'
    if os.path.exists(fn): return pickle.load(open(fn,'rb'))
    if which=='c2': w=synth.make_workload(2,B=cnt)
    else: w=synth.make_sweep(int(which[1:]), states_per_point=1, max_points=cnt)
    inst=instances(w,cnt); pickle.dump(inst,open(fn,'wb')); return inst
if __name__=="__main__":
    which=sys.argv[1]; cnt=int(sys.argv[2]); kw=eval("dict(%s)"%(sys.argv[3] if len(sys.argv)>3 else ""))
    inst=get(which,cnt)
    its=[];errs=[];oks=[]
    for p,r,cq in inst:
        u,it,ok=CM.ipm_solve(cq,**kw); its.append(it); errs.append(ctrl_err(cq,r,u)); oks.append(ok)
    its=np.array(its); errs=np.array(errs)
    print(which,kw,"iters mean %.2f p99 %d max %d | err max %.2e | not converged %d | >1e-4 %d"%(its.mean(),np.percentile(its,99),its.max(),errs.max(),len(oks)-sum(oks),(errs>1e-4).sum()))
