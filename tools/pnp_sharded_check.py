"""torchrun check of the hypothesis-sharded PnP path (NCCL, device-resident key): every rank must return the single-GPU answer."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
import ransac_b200
from ransac_b200 import dist as rdist, synth
rank, local = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local); device = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=device)
P, px, _ = synth.pnp_set(20000, 0.5, np.random.default_rng(3))
ctx = ransac_b200.Context(local)
ok, r, t, inl, info = rdist.solve_pnp_ransac_sharded(ctx, P, px, synth.K_1898, 8.0, 16384, seed=9, arith=ransac_b200.ARITH_FAST, device=device, solver=ransac_b200.SOLVER_FAST)
ok1, r1, t1, inl1, info1 = ctx.solve_pnp_ransac(P, px, synth.K_1898, 16384, 8.0, 0.99, sampler=ransac_b200.SAMPLER_PHILOX, seed=9, arith=ransac_b200.ARITH_FAST, solver=ransac_b200.SOLVER_FAST)
assert ok and ok1 and np.array_equal(inl, inl1) and np.array_equal(r, r1) and np.array_equal(t, t1), (rank, len(inl), len(inl1))
src, dst, _ = synth.homography_set(20000, 0.5, np.random.default_rng(4))
H, m, i = rdist.find_homography_sharded(ctx, src, dst, 3.0, 16384, seed=9, arith=ransac_b200.ARITH_FAST, solver=ransac_b200.SOLVER_FAST, device=device)
H1, m1, i1 = ctx.find_homography(src, dst, 3.0, max_iters=16384, sampler=ransac_b200.SAMPLER_PHILOX, seed=9, arith=ransac_b200.ARITH_FAST, solver=ransac_b200.SOLVER_FAST)
assert np.array_equal(H, H1) and np.array_equal(m, m1)
dist.barrier(); print("sharded == single ok", rank, len(inl), int(m.sum()), flush=True)
dist.destroy_process_group()
