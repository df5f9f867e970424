"""Multi-GPU sharding of the hot path; one process per GPU, torch.distributed.

Two modes, as in BASELINE.json's north star:

* pairs of a registration set are split across ranks with no collective during compute
  (`shard_pairs`, the reference's DistributedSampler split, Experiments/dataloader/data_loaders.py:111-116),
  followed by one gather of the per-pair stat rows (`gather_rows`; the reference writes per-rank
  .npy files instead, Experiments/test.py:257);
* the hypotheses of ONE pair are split across ranks (`ransac_rigid_sharded`): a hypothesis is a
  pure function of (seed, id), rank g scores a contiguous slice of every round, and the ranks
  exchange one 8-byte packed (inlier count, hypothesis id) key per round -- inside the kernel that
  ends the round, through peer mailboxes over NVLink (the library's own communicator, `init_comm`), or
  with a torch.distributed MAX all-reduce.  Ties go to the lowest id, so the result is independent of
  the number of GPUs; every rank regenerates the winning model from the id, so nothing is broadcast.

`backend` is the object that does the scoring: lidarregistration_b200.engine (CUDA) by default.
The CPU tests of this module (gloo, world size 2) pass a stand-in built on the oracle.
"""
import numpy as np
import torch
import torch.distributed as dist


def world(group=None):
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(group), dist.get_world_size(group)
    return 0, 1


def shard_pairs(num_pairs, rank, world_size):
    """pair p -> rank p mod G; no padding / duplication (unlike DistributedSampler's drop_last=False)."""
    return list(range(rank, num_pairs, world_size))


def ransac_rigid_pairs(pairs, params, backend=None, group=None):
    """RANSAC over this rank's share of a registration set: `pairs` is the WHOLE list of (src, tgt)
    correspondence sets (or callables producing them); rank r runs pairs r, r + G, ... through one batched call
    (engine.ransac_rigid_batch: two pairs in flight, no host round trip between pairs) and no collective.
    Returns [(pair index, result dict), ...] for the pairs of this rank."""
    if backend is None:
        from . import engine as backend
    rank, ws = world(group)
    mine = shard_pairs(len(pairs), rank, ws)
    data = [pairs[i]() if callable(pairs[i]) else pairs[i] for i in mine]
    return list(zip(mine, backend.ransac_rigid_batch(data, params)))


def shard_range(lo, hi, rank, world_size):
    """contiguous, near-equal slice of the hypothesis ids [lo, hi) owned by `rank`"""
    n = hi - lo
    a = lo + (n * rank) // world_size
    b = lo + (n * (rank + 1)) // world_size
    return a, b


def gather_rows(rows, group=None):
    """per-rank [k_r, C] float arrays -> one [sum k_r, C] array on every rank (rank order)"""
    rank, ws = world(group)
    rows = np.asarray(rows, dtype=np.float64)
    if ws == 1:
        return rows
    out = [None] * ws
    dist.all_gather_object(out, rows, group=group)
    return np.concatenate([o for o in out if len(o)], axis=0) if any(len(o) for o in out) else rows


def init_comm(group=None, backend=None):
    """Connect the library-owned communicator of hypothesis sharding (lr_comm_init / lr_comm_connect): every rank
    allocates its mailbox, the 64-byte IPC handles travel through torch.distributed, every rank maps its peers'
    mailboxes (NVLink / NVSwitch peer memory).  Returns True when the fused transport is ready on this rank set,
    False when it cannot be set up (then ransac_rigid_sharded keeps using the all-reduce transport)."""
    if backend is None:
        from . import engine as backend
    rank, ws = world(group)
    if not hasattr(backend, "comm_init"):
        return False
    r, w = backend.comm_world()
    if w == ws and r == rank:
        return True
    if w != 0:
        backend.comm_destroy()

    def gather(b):
        if ws == 1:
            return [b]
        out = [None] * ws
        dist.all_gather_object(out, b, group=group)
        return out

    ok = 1
    try:
        backend.comm_init(rank, ws, gather)
    except RuntimeError:
        ok = 0
    if ws > 1:  # all ranks or none
        flag = torch.tensor([ok], dtype=torch.int32, device=torch.device("cuda", torch.cuda.current_device())
                            if dist.get_backend(group) == "nccl" else "cpu")
        dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=group)
        ok = int(flag.item())
    if not ok:
        try:
            backend.comm_destroy()
        except RuntimeError:
            pass
    return bool(ok)


def ransac_rigid_sharded(src, tgt, params, backend=None, group=None, device=None, transport="auto", want_mask=False):
    """RANSAC over one correspondence set with the hypotheses of every round split across ranks.

    params: engine.LrRansacParams.  Returns the same dict as engine.ransac_rigid on every rank.
    Semantics are those of the single-GPU call: with a confidence < 1 the exit is evaluated at
    round ends (round_size hypotheses, all ranks together); with a fixed budget the whole budget is
    one round per rank and a single exchange.  Count scoring only: the exchanged key packs (inlier count,
    hypothesis id) into 64 bits, a quantised MSAC score needs up to 47 bits on its own (shard MSAC runs by pair).

    transport: "p2p" = the library's own communicator (init_comm): the key exchange is part of the kernel that
    ends a round, peers write each other's mailboxes over NVLink, one host synchronisation per call;
    "allreduce" = lr_ransac_shard + torch.distributed all_reduce(MAX) + lr_ransac_finalize (any backend, also
    what the CPU tests run over gloo); "auto" = p2p when a communicator is connected, else allreduce.
    """
    if backend is None:
        from . import engine as backend
    if int(getattr(params, "scoring", 0)) != 0:
        raise ValueError("ransac_rigid_sharded: hypothesis sharding supports count scoring only "
                         "(LR_SCORE_COUNT); use ransac_rigid_pairs for LR_SCORE_MSAC runs")
    rank, ws = world(group)
    if transport == "auto":
        transport = "allreduce"
        if hasattr(backend, "comm_world"):
            r, w = backend.comm_world()
            if w == ws and r == rank:
                transport = "p2p"
    if transport == "p2p":
        out = backend.ransac_rigid_sharded(src, tgt, params, want_mask=want_mask)
        out["transport"] = "p2p"
        return out
    if hasattr(backend, "to_dev_f32"):  # any dtype / layout / host array in, contiguous fp32 device tensors down
        src, tgt = backend.to_dev_f32(src), backend.to_dev_f32(tgt)
    n = int(src.shape[0])
    m = int(params.sample_size)
    max_iters = int(params.max_iters)
    conf = float(params.confidence)
    use_conf = conf < 1.0
    R = int(params.round_size) if use_conf else max(max_iters, 1)
    if n < m:  # Open3D: |corres| < ransac_n -> identity (App. B); same as engine.ransac_rigid
        out = dict(T=np.eye(4), T_refit=np.eye(4), mask=None, iters_run=0, n_scored=0, n_rechecked=0, best_id=-1,
                   best_count=-1, refit_count=0, transport="allreduce")
        return out
    if device is None:
        device = src.device if torch.is_tensor(src) else torch.device("cpu")
    key = torch.zeros(1, dtype=torch.int64, device=device)
    done = 0
    while done < max_iters:
        lo, hi = done, min(done + R, max_iters)
        a, b = shard_range(lo, hi, rank, ws)
        if b > a:
            backend.ransac_shard(src, tgt, params, a, b, key)
        if ws > 1:
            dist.all_reduce(key, op=dist.ReduceOp.MAX, group=group)
        done = hi
        if use_conf:
            cnt, _ = backend.key_unpack(int(key.item()))  # one 8-byte D2H per round
            if cnt > 0 and done >= backend.conf_iters(cnt, n, m, conf, max_iters):
                break
    out = backend.ransac_finalize(src, tgt, params, int(key.item()))
    out["iters_run"] = done
    out["transport"] = "allreduce"
    return out
