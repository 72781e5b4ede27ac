"""Two or more ranks (torchrun, NCCL): every rank solves its shard of one global batch on its own GPU, the per-rank results
are gathered to rank 0 with lbmpc_b200.dist.gather_results (NCCL gather) and the statistics all-reduced; rank 0 checks the
gathered arrays bit for bit against a single-GPU solve of the whole batch and against the CPU oracle.
usage: torchrun --nproc-per-node W tools/nccl_gather_check.py [batch]"""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(ROOT, "learning-based-mpc_b200"), os.path.join(ROOT, "oracle")):
    sys.path.insert(0, p)
import numpy as np
import torch
import torch.distributed as dist
import lbmpc_b200
from lbmpc_b200.dist import gather_results, reduce_stats, sample_initial_states, shard_range

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
nb = int(sys.argv[1]) if len(sys.argv) > 1 else 4099          # not divisible by the world size: ragged shards
N = 50
mdl = lbmpc_b200.moore_greitzer_model("LBMPC")
X0 = sample_initial_states(nb, seed=5)
lo, hi = shard_range(nb, rank, world)
sol = lbmpc_b200.Solver(mdl, "C", "LBMPC", N, device=local, device_pointers=True, kernel="warp")
o = sol.solve_batch(torch.from_numpy(X0[lo:hi]).to(dev), want_x=False)
g = gather_results({"u0": o["uc"][:, 0, 0].contiguous(), "obj": o["obj"], "iters": o["iters"], "status": o["status"]}, nb, dst=0)
st = reduce_stats(o["status"], o["iters"], o["obj"])
ok = True
if rank == 0:
    whole = sol.solve_batch(torch.from_numpy(X0).to(dev), want_x=False)
    for k, ref in (("u0", whole["uc"][:, 0, 0]), ("obj", whole["obj"]), ("iters", whole["iters"]), ("status", whole["status"])):
        ok &= bool(torch.equal(g[k], ref))
    from oracle_py import OracleProblem
    r = OracleProblem("C", "LBMPC", mdl, N).solve_batch(X0[:256], nthreads=8)
    ok &= bool(np.array_equal(r["status"], g["status"][:256].cpu().numpy()))
    ok &= bool(np.abs(r["iters"] - g["iters"][:256].cpu().numpy()).max() <= 1)
    ok &= st["n_optimal"] + st["n_maxiter"] + st["n_infeasible"] + st["n_numerical"] == nb
    ok &= st["sum_iters"] == int(whole["iters"].sum())
    print(json.dumps({"nccl_gather_check": "ok" if ok else "MISMATCH", "world": world, "batch": nb, "backend": dist.get_backend(), "stats": st}), flush=True)
dist.barrier()
dist.destroy_process_group()
sys.exit(0 if ok else 1)
