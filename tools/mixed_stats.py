import sys, os, numpy as np
ROOT='/root/repo'
sys.path.insert(0, os.path.join(ROOT, "learning-based-mpc_b200")); sys.path.insert(0, os.path.join(ROOT, "oracle")); sys.path.insert(0, os.path.join(ROOT,"tests"))
import lbmpc_b200
from lbmpc_b200.dist import sample_initial_states
from oracle_py import OracleProblem
mdl = lbmpc_b200.moore_greitzer_model("LBMPC")
X0 = sample_initial_states(2048, 77)
ref = OracleProblem("C", "LBMPC", mdl, 50).solve_batch(X0, nthreads=16)
got = lbmpc_b200.Solver(mdl, "C", "LBMPC", 50, max_batch=2048, kernel="mixed").solve_batch(X0)
print("status mismatch", (got["status"]!=ref["status"]).sum(), np.bincount(got["status"],minlength=4), np.bincount(ref["status"],minlength=4))
d = got["iters"].astype(int)-ref["iters"].astype(int)
print("dit hist", {int(k):int(v) for k,v in zip(*np.unique(d, return_counts=True))})
ok=(ref["status"]==0)&(got["status"]==0)
e=np.abs(got["uc"][ok].reshape(ok.sum(),-1)-ref["uc"][ok].reshape(ok.sum(),-1)).max(1)/np.maximum(1,np.abs(ref["uc"][ok].reshape(ok.sum(),-1)).max(1))
print("u err <1e-8",(e<1e-8).mean(),"<1e-6",(e<1e-6).mean(),"<1e-4",(e<1e-4).mean(),"max",e.max())
print("obj", (np.abs(got["obj"][ok]-ref["obj"][ok])/np.maximum(1,np.abs(ref["obj"][ok]))).max())
bad=np.nonzero(np.abs(d)>3)[0]; print("bad", bad[:10], got["iters"][bad[:10]], ref["iters"][bad[:10]], got["status"][bad[:10]], ref["status"][bad[:10]])
