"""Host-side logic of the N > 1 paths on CPU: world size 2 over gloo.  The scoring backend is a
stand-in built on the oracle (test infrastructure); the product default is the CUDA engine."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from lidarregistration_b200 import engine, parallel, synthetic
from oracle import lr_oracle as O


class OracleBackend:
    """what engine.ransac_shard / ransac_finalize / conf_iters do, on the CPU oracle"""
    key_unpack = staticmethod(engine.key_unpack)
    conf_iters = staticmethod(O.conf_iters)

    @staticmethod
    def ransac_shard(src, tgt, p, lo, hi, key):
        samples = np.stack([O.sample(p.seed, h, p.sampler, p.sample_size, len(src))
                            for h in range(lo, hi)]).astype(np.int32)
        counts, best = O.score_samples(src, tgt, samples, p.threshold, bool(p.use_elc), p.elc_ratio)
        if best >= 0:
            key[0] = max(int(key[0]), engine.key_pack(int(counts[best]), lo + best))

    @staticmethod
    def ransac_finalize(src, tgt, p, key):
        cnt, hid = engine.key_unpack(key)
        T = np.eye(4)
        if cnt > 0:
            s = O.sample(p.seed, hid, p.sampler, p.sample_size, len(src))
            T = O.kabsch(src[s].astype(float), tgt[s].astype(float))
        return dict(T=T, best_id=hid, best_count=cnt)

    @staticmethod
    def ransac_rigid_batch(pairs, p):
        out = []
        for src, tgt in pairs:
            if p.scoring == engine.SCORE_MSAC:  # GC semantics shard by pair like any other run
                r = O.ransac_gc(src, tgt, m=p.sample_size, sampler=p.sampler, use_elc=bool(p.use_elc), thr=p.threshold,
                                conf=p.confidence, max_iters=p.max_iters, round_size=p.round_size, seed=p.seed,
                                lo_rounds=p.lo_rounds, lo_trials=p.lo_trials, lsq_iters=p.lsq_iters)
                out.append(dict(T=r["T"], best_id=r["best_id"], best_count=r["best_inliers"],
                                final_score=r["final_score"]))
                continue
            r = O.ransac(src, tgt, m=p.sample_size, sampler=p.sampler, use_elc=bool(p.use_elc), thr=p.threshold,
                         conf=p.confidence, max_iters=p.max_iters, round_size=p.round_size, seed=p.seed)
            out.append(dict(T=r["T"], best_id=r["best_id"], best_count=r["best_count"]))
        return out


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    d = synthetic.make_correspondences(1500, inlier_ratio=0.35, seed=123)
    out = {}
    for name, conf, iters, R in (("fixed", 1.0, 3000, 65536), ("conf", 0.999, 20000, 512)):
        p = engine.make_params(confidence=conf, max_iters=iters, seed=7, round_size=R, use_elc=True)
        r = parallel.ransac_rigid_sharded(d["src"], d["tgt"], p, backend=OracleBackend)
        out[name] = (r["best_id"], r["best_count"], r["iters_run"], r["T"].tolist())
    rows = np.full((3 + rank, 22), float(rank))
    out["rows"] = parallel.gather_rows(rows).tolist()
    out["pairs"] = parallel.shard_pairs(11, rank, world)
    # pairs of a set sharded by rank, no collective: every pair is run exactly once, by rank p mod G
    sets = [synthetic.make_correspondences(400 + 50 * k, inlier_ratio=0.4, seed=900 + k) for k in range(5)]
    p = engine.make_params(confidence=1.0, max_iters=500, seed=3, use_elc=True)
    mine = parallel.ransac_rigid_pairs([(d["src"], d["tgt"]) for d in sets], p, backend=OracleBackend)
    out["set"] = [(i, r["best_id"], r["best_count"]) for i, r in mine]
    pg = engine.make_params(confidence=1.0, max_iters=500, seed=3, use_elc=True, scoring=engine.SCORE_MSAC,
                            lo_rounds=3, lo_trials=6, lsq_iters=2)
    mine = parallel.ransac_rigid_pairs([(d["src"], d["tgt"]) for d in sets], pg, backend=OracleBackend)
    out["set_gc"] = [(i, r["best_id"], r["final_score"]) for i, r in mine]
    try:
        parallel.ransac_rigid_sharded(sets[0]["src"], sets[0]["tgt"], pg, backend=OracleBackend)
        out["gc_sharded_refused"] = False
    except ValueError:
        out["gc_sharded_refused"] = True
    q.put((rank, out))
    dist.barrier()
    dist.destroy_process_group()


def test_sharding_helpers():
    assert parallel.shard_pairs(7, 0, 2) == [0, 2, 4, 6] and parallel.shard_pairs(7, 1, 2) == [1, 3, 5]
    cover = []
    for r in range(3):
        a, b = parallel.shard_range(100, 1111, r, 3)
        cover += list(range(a, b))
    assert cover == list(range(100, 1111))
    assert parallel.shard_range(5, 5, 0, 2) == (5, 5)


def test_world2_matches_single_process():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=300) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    d = synthetic.make_correspondences(1500, inlier_ratio=0.35, seed=123)
    for name, conf, iters, R in (("fixed", 1.0, 3000, 65536), ("conf", 0.999, 20000, 512)):
        ref = O.ransac(d["src"], d["tgt"], conf=conf, max_iters=iters, round_size=R, seed=7)
        for rank in (0, 1):
            bid, bcnt, run, T = res[rank][name]
            assert (bid, bcnt, run) == (ref["best_id"], ref["best_count"], ref["iters_run"]), (name, rank)
            assert np.array_equal(np.array(T), ref["T"])
    rows = np.array(res[0]["rows"])
    assert rows.shape == (7, 22) and np.all(rows[:3] == 0) and np.all(rows[3:] == 1)
    assert res[0]["rows"] == res[1]["rows"]
    assert sorted(res[0]["pairs"] + res[1]["pairs"]) == list(range(11))
    got = sorted(res[0]["set"] + res[1]["set"])
    assert [g[0] for g in got] == list(range(5)) and [g[0] for g in res[1]["set"]] == [1, 3]
    for k, bid, bcnt in got:
        d = synthetic.make_correspondences(400 + 50 * k, inlier_ratio=0.4, seed=900 + k)
        ref = O.ransac(d["src"], d["tgt"], conf=1.0, max_iters=500, seed=3)
        assert (bid, bcnt) == (ref["best_id"], ref["best_count"]), k
    # GC semantics (MSAC + LO + least squares) shard by pair the same way; hypothesis sharding refuses them
    got = sorted(res[0]["set_gc"] + res[1]["set_gc"])
    assert [g[0] for g in got] == list(range(5)) and res[0]["gc_sharded_refused"] and res[1]["gc_sharded_refused"]
    for k, bid, fq in got:
        d = synthetic.make_correspondences(400 + 50 * k, inlier_ratio=0.4, seed=900 + k)
        ref = O.ransac_gc(d["src"], d["tgt"], conf=1.0, max_iters=500, seed=3, lo_rounds=3, lo_trials=6, lsq_iters=2)
        assert (bid, fq) == (ref["best_id"], ref["final_score"]), k
