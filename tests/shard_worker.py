"""Worker of tests/test_gpu_sharded.py: one process per rank (torchrun), checks that the hypothesis-sharded run
(library communicator over peer memory, and the all-reduce transport) returns exactly what the single-GPU call
returns.  Ranks share a GPU when there are fewer GPUs than ranks (functional check only: the contexts time-slice)."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from lidarregistration_b200 import engine, parallel, synthetic  # noqa: E402


def same(a, b, refit_exact=True):
    # (the all-reduce transport exchanges the key only: its n_scored is not summed over the ranks)
    for k in ("best_id", "best_count", "iters_run", "refit_count") + (("n_scored",) if refit_exact else ()):
        assert a[k] == b[k], (k, a[k], b[k])
    assert np.array_equal(a["T"], b["T"])
    if refit_exact:
        assert np.array_equal(a["T_refit"], b["T_refit"])
    else:
        assert np.abs(a["T_refit"] - b["T_refit"]).max() < 1e-9


def checks(rank, ws, transports):
    out = []
    cases = [
        dict(n=12000, ratio=0.3, iters=200000, conf=1.0, elc=True, m=3, sampler=engine.SAMPLER_UNIFORM, rs=65536),
        dict(n=12000, ratio=0.3, iters=400000, conf=0.9995, elc=True, m=3, sampler=engine.SAMPLER_UNIFORM, rs=16384),
        dict(n=5000, ratio=0.4, iters=20000, conf=1.0, elc=False, m=3, sampler=engine.SAMPLER_UNIFORM, rs=65536),
        dict(n=8000, ratio=0.5, iters=100000, conf=0.9995, elc=True, m=4, sampler=engine.SAMPLER_REPLACE, rs=8192),
        dict(n=8000, ratio=0.3, iters=150000, conf=1.0, elc=True, m=3, sampler=engine.SAMPLER_PROSAC, rs=65536),
        dict(n=2, ratio=1.0, iters=1000, conf=1.0, elc=True, m=3, sampler=engine.SAMPLER_UNIFORM, rs=65536),
        dict(n=3000, ratio=0.3, iters=5, conf=1.0, elc=False, m=3, sampler=engine.SAMPLER_UNIFORM, rs=65536),
    ]
    for ci, c in enumerate(cases):
        d = synthetic.make_correspondences(c["n"], inlier_ratio=c["ratio"], seed=900 + ci)
        p = engine.make_params(confidence=c["conf"], max_iters=c["iters"], seed=5 + ci, use_elc=c["elc"], sample_size=c["m"],
                               sampler=c["sampler"], round_size=c["rs"])
        src, tgt = torch.from_numpy(d["src"]).cuda(), torch.from_numpy(d["tgt"]).cuda()
        single = engine.ransac_rigid(src, tgt, p, want_mask=True)
        for transport in transports:
            for rep in range(2):
                r = parallel.ransac_rigid_sharded(src, tgt, p, transport=transport, want_mask=(transport == "p2p"))
                assert r["transport"] == transport
                if c["n"] >= c["m"]:
                    same(r, single, refit_exact=(transport == "p2p"))
                    if transport == "p2p":
                        assert torch.equal(r["mask"], single["mask"])
                else:
                    assert np.array_equal(r["T"], np.eye(4)) and r["best_id"] == -1
        out.append((ci, single["best_id"], single["best_count"], single["iters_run"]))
    return out


def main():
    rank, ws = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    ngpu = torch.cuda.device_count()
    shared_gpu = ngpu < ws
    torch.cuda.set_device(local % ngpu)
    if shared_gpu:
        dist.init_process_group("gloo")
    else:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local % ngpu))
    ok = parallel.init_comm()
    if not ok:
        if rank == 0:
            print("SHARD_SKIP communicator could not be connected (cudaIpc unavailable?)", flush=True)
        dist.destroy_process_group()
        return
    # over gloo (ranks sharing one GPU) the all-reduce transport would need CPU tensors in the collective: the
    # shared-GPU run checks the peer-mailbox transport only
    transports = ("p2p",) if shared_gpu else ("p2p", "allreduce")
    res = checks(rank, ws, transports)
    dist.barrier()
    if rank == 0:
        print("SHARD_OK world=%d shared_gpu=%d cases=%s" % (ws, int(shared_gpu), res), flush=True)
    engine.comm_destroy()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
