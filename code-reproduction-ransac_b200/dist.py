"""Multi-GPU form of the hot path: one process per GPU, hypotheses sharded by id, one 8-byte MAX all-reduce.

SURVEY.md §8(e): the points are replicated on every rank (<= 16 MB), rank r scores the Philox hypothesis ids
[r*H, (r+1)*H) — the sampler is counter-based, so the partition does not change any sample — and reduces them to
one packed key per problem, (count << 32) | (0xFFFFFFFF - id).  MAX over ranks picks the best count with the
lowest id on ties (OpenCV's "first strictly better" rule).  Every rank then re-derives the winning sample from
its id and runs the finalize kernel (mask, refit, LM) redundantly, so no broadcast is needed.

torch.distributed is plumbing here (process group + the NCCL collective); on CPU-only test runs the same code
runs over gloo with the reduction applied to keys produced by the caller."""
import numpy as np

from . import api


def reduce_keys_max(keys, group=None, device=None):
    """MAX all-reduce of uint64 keys across the process group (NCCL on `device`, gloo on CPU).

    The keys are < 2^63 (counts are < 2^31), so they travel as int64 — NCCL/gloo have no uint64 MAX."""
    import torch
    import torch.distributed as dist
    keys = np.ascontiguousarray(np.asarray(keys, dtype=np.uint64))
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return keys.copy()
    if int(keys.max(initial=0)) >= 2 ** 63:
        raise ValueError("key overflow")
    t = torch.from_numpy(keys.view(np.int64).copy())
    if device is not None:
        t = t.to(device, non_blocking=False)
    dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    return t.cpu().numpy().view(np.uint64)


_KEYS = {}


def _device_keys(Q, device):
    """A cached int64 device tensor of Q packed keys (the all-reduce buffer)."""
    import torch
    k = (str(device), int(Q))
    if k not in _KEYS:
        _KEYS[k] = torch.zeros(int(Q), dtype=torch.int64, device=device)
    return _KEYS[k]


def shard_range(total, rank, world):
    """Contiguous split of `total` hypothesis ids: rank r gets [begin, begin+count)."""
    base, rem = divmod(int(total), int(world))
    begin = rank * base + min(rank, rem)
    return begin, base + (1 if rank < rem else 0)


def run_sharded(problem, thr, hyp_begin, hyp_count, seed=0, arith=api.ARITH_EXACT, confidence=0.995,
                mask_semantics=api.MASK_CV413, refine=True, solver=api.SOLVER_EXACT, group=None, device=None):
    """Score this rank's shard of a resident problem, reduce, finish.  Results stay on the device (problem.fetch())."""
    p = api.make_params(thr, hyp_count, confidence, sampler=api.SAMPLER_PHILOX, seed=seed, arith=arith,
                        mask_semantics=mask_semantics, refine=refine, hyp_begin=hyp_begin, solver=solver)
    import torch.distributed as dist
    multi = dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1
    if device is not None and (not multi or dist.get_backend(group) == "nccl"):
        # keys never leave the GPU: score -> NCCL MAX all-reduce -> finish, all enqueued on the library's stream
        import torch
        stream = torch.cuda.ExternalStream(problem.ctx.stream, device=device)
        with torch.cuda.stream(stream):
            t = _device_keys(problem.Q, device)
            problem.score_shard_dev(p, t.data_ptr())
            if multi:
                dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)     # keys < 2^63: int64 MAX == uint64 MAX
            problem.finish_dev(p, t.data_ptr())
        return None
    keys = problem.score_shard(p)
    best = reduce_keys_max(keys, group=group, device=device)
    problem.finish(p, best)
    return best


def find_homography_sharded(ctx, src, dst, thr, total_hypotheses, seed=0, arith=api.ARITH_EXACT, group=None,
                            device=None, problem=None, **kw):
    """End-to-end sharded call: host points in, (H, mask, info) out on every rank; hypothesis ids
    [0, total_hypotheses) are split over the ranks of `group`."""
    import torch.distributed as dist
    rank, world = (dist.get_rank(group), dist.get_world_size(group)) if dist.is_available() and dist.is_initialized() else (0, 1)
    begin, count = shard_range(total_hypotheses, rank, world)
    # `problem`: a handle from an earlier call whose device buffers are reused (no cudaMalloc per call)
    prob = problem
    if prob is None:
        prob = ctx.upload(src, dst)
    else:
        prob.reupload(src, dst)
    try:
        run_sharded(prob, thr, begin, count, seed=seed, arith=arith, group=group, device=device, **kw)
        H, mask, infos = prob.fetch()
    finally:
        if problem is None:
            prob.free()
    ok = infos[0]["status"] == api.OK
    return (H[0] if ok else None), mask[0].reshape(-1, 1), infos[0]


# ---- PnP path (cv2.solvePnPRansac, main_v1.py:497): same sharding scheme ---------------------------------------------
def run_sharded_pnp(problem, thr, hyp_begin, hyp_count, seed=0, arith=api.ARITH_EXACT, confidence=0.99, refine=True, group=None,
                    device=None, solver=api.SOLVER_EXACT):
    """Score this rank's hypothesis-id shard of a resident PnPProblem, MAX-reduce the packed keys, finish on every rank."""
    p = api.make_p_params(thr, hyp_count, confidence, sampler=api.SAMPLER_PHILOX, seed=seed, arith=arith, refine=refine,
                          hyp_begin=hyp_begin, solver=solver)
    import torch.distributed as dist
    if device is not None and dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1 \
            and dist.get_backend(group) == "nccl":
        import torch
        stream = torch.cuda.ExternalStream(problem.ctx.stream, device=device)
        with torch.cuda.stream(stream):
            t = _device_keys(problem.Q, device)
            problem.score_shard_dev(p, t.data_ptr())
            dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
            problem.finish_dev(p, t.data_ptr())
        return None
    keys = problem.score_shard(p)
    best = reduce_keys_max(keys, group=group, device=device)
    problem.finish(p, best)
    return best


def solve_pnp_ransac_sharded(ctx, obj, img, K, thr, total_hypotheses, seed=0, arith=api.ARITH_EXACT, group=None, device=None,
                             problem=None, **kw):
    """End-to-end sharded cv2.solvePnPRansac: host points in, (ok, rvec, tvec, inliers, info) out on every rank."""
    import torch.distributed as dist
    rank, world = (dist.get_rank(group), dist.get_world_size(group)) if dist.is_available() and dist.is_initialized() else (0, 1)
    begin, count = shard_range(total_hypotheses, rank, world)
    prob = problem
    if prob is None:
        prob = ctx.upload_pnp(obj, img, K)
    else:
        prob.reupload(obj, img, K)
    try:
        run_sharded_pnp(prob, thr, begin, count, seed=seed, arith=arith, group=group, device=device, **kw)
        rvec, tvec, inliers, infos = prob.fetch()
    finally:
        if problem is None:
            prob.free()
    ok = infos[0]["status"] == api.OK
    return ok, rvec[0].reshape(3, 1), tvec[0].reshape(3, 1), (inliers[0].reshape(-1, 1) if ok else None), infos[0]


# ---- batched calls shard by PROBLEM: no exchange on the data path, one all-gather of the per-problem scores -----------
def allgather_rows(local_rows, total, group=None, device=None):
    """All ranks contribute their contiguous block of rows (shard_range order) of a (total, k) float64 table; every rank
    gets the whole table.  NCCL on `device`, gloo on CPU.  (Q x (score, id), 64 KB at Q = 4096.)"""
    import torch
    import torch.distributed as dist
    local_rows = np.ascontiguousarray(np.asarray(local_rows, dtype=np.float64))
    k = local_rows.shape[1] if local_rows.ndim == 2 else 1
    local_rows = local_rows.reshape(-1, k)
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        assert len(local_rows) == total
        return local_rows.copy()
    world = dist.get_world_size(group)
    counts = [shard_range(total, r, world)[1] for r in range(world)]
    pad = max(counts)
    buf = np.zeros((pad, k))
    buf[:len(local_rows)] = local_rows
    t = torch.from_numpy(buf)
    if device is not None:
        t = t.to(device)
    out = [torch.empty_like(t) for _ in range(world)]
    dist.all_gather(out, t, group=group)
    return np.concatenate([o.cpu().numpy()[:c] for o, c in zip(out, counts)], axis=0)


def camera_sweep_sharded(ctx, pos3d, pixels, cams, thr, group=None, device=None, **kw):
    """find_homographies + arg-min (main_v1.py:254-297, :863-866) with the candidate cameras sharded over the ranks: each
    rank runs the fused device sweep on its block of candidates; the (err1, err2) rows are all-gathered and every rank
    takes the same arg-min.  Returns (scores (Q,2), best index)."""
    import torch.distributed as dist
    rank, world = (dist.get_rank(group), dist.get_world_size(group)) if dist.is_available() and dist.is_initialized() else (0, 1)
    cams = np.asarray(cams, dtype=np.float64).reshape(-1, 3)
    Q = len(cams)
    begin, count = shard_range(Q, rank, world)
    local = ctx.camera_sweep(pos3d, pixels, cams[begin:begin + count], thr, **kw)["scores"] if count > 0 else np.zeros((0, 2))
    scores = allgather_rows(local, Q, group=group, device=device)
    err2 = scores[:, 1].copy()
    err2[err2 == 0] = 1000000                      # main_v1.py:865
    return scores, int(np.argmin(err2))
