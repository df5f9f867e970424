#!/usr/bin/env python
"""bench.py -- the hot path's headline benchmark (contract: the build prompt, section 4).

metric : pairs/sec @1M RANSAC iters (+ MNN match ms as `mnn_match`), BASELINE.json
step   : PAIRS_PER_STEP synthetic cfg-3 pairs (30k correspondences, 30 % inliers), each run for
         1,000,000 hypotheses (3-point sample, edge-length pre-rejection, Kabsch, inlier count,
         least-squares refit) -- the reference's `--algo RANSAC --iters 1000000` with its default
         `--fast_rejection ELC`, confidence exit disabled so the whole budget is spent
value  : device-resident inputs, whole job over all ranks (pairs sharded by rank, no collective)
e2e    : the same work through the reference-facing call findRigidTransform(...) with HOST
         buffers (pinned), H2D + D2H inside the timed region
--impl reference : the CPU restatement of the reference's path (oracle/) on all host cores.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_CORR = 30000
INLIER_RATIO = 0.3
ITERS = 1000000
THRESH = 0.6
CFG_SEED = 51 + 3000
PAIRS_PER_STEP = 16  # (4 in round 1: a 1.3 ms step left the 8-rank run at the mercy of host launch jitter)
MATCH_N = 50000
FLOPS_PER_TEST = 27  # 12 FMA + 3 for R p + t - q and |.|^2, compare excluded
METRIC = "pairs/sec @1M RANSAC iters"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-elc", action="store_true", help="headline regime without edge-length rejection")
    ap.add_argument("--skip-extras", action="store_true", help="only the headline line (used under ncu)")
    ap.add_argument("--mode", default="pairs", choices=["pairs", "hyp"],
                    help="pairs: the rank's pairs, no collective (weak scaling, the default headline); hyp: the hypotheses "
                         "of ONE pair split over the ranks (BASELINE cfg 3 as worded, strong scaling) as the headline")
    return ap.parse_args()


def config_dict(use_elc, pairs_per_step=PAIRS_PER_STEP):
    return {"workload": "cfg3: RANSAC --iters 1000000 on one 30k-correspondence synthetic pair, 30%% inliers, "
                        "3-point samples, ELC %s, fixed budget (confidence exit off), count scoring + LSQ refit"
                        % ("on (reference default --fast_rejection ELC)" if use_elc else "off (every hypothesis scored)"),
            "n_correspondences": N_CORR, "iters": ITERS, "inlier_ratio": INLIER_RATIO, "threshold_m": THRESH,
            "pairs_per_step": pairs_per_step, "elc": bool(use_elc),
            "l2": "a 512 MB buffer is written between timed steps (L2 flush); inputs are 0.72 MB per pair"}


def make_pairs(count, base_seed):
    from lidarregistration_b200 import synthetic
    return [synthetic.make_correspondences(N_CORR, INLIER_RATIO, seed=base_seed + p) for p in range(count)]


# ----------------------------------------------------------------------------- reference arm
def cpu_ransac_rate(pairs, use_elc, threads=None):
    """oracle (CPU restatement, OpenMP) on whole pairs at the full budget -> (pairs/s, cores, seconds)"""
    from oracle import lr_oracle as O
    if threads:
        O.set_threads(threads)
    t0 = time.perf_counter()
    for d in pairs:
        O.ransac(d["src"], d["tgt"], m=3, sampler=0, use_elc=use_elc, thr=THRESH, conf=1.0, max_iters=ITERS,
                 round_size=65536, seed=51, refit=True)
    dt = time.perf_counter() - t0
    return len(pairs) / dt, O.num_threads(), dt


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    use_elc = not args.no_elc
    pairs = make_pairs(PAIRS_PER_STEP, CFG_SEED)  # the same pairs rank 0 of the CUDA arm runs
    from oracle import lr_oracle as O
    # all the host threads this process may use (torchrun exports OMP_NUM_THREADS=1 for its workers)
    try:
        O.set_threads(len(os.sched_getaffinity(0)))
    except (AttributeError, OSError):
        O.set_threads(os.cpu_count() or 1)
    cores = O.num_threads()
    for _ in range(args.warmup):
        cpu_ransac_rate(pairs, use_elc)
    times = []
    for _ in range(args.steps):
        _, _, dt = cpu_ransac_rate(pairs, use_elc)
        times.append(dt)
    ms = 1e3 * sum(times) / len(times)
    value = PAIRS_PER_STEP * 1e3 / ms
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": "pairs/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": config_dict(use_elc),
            "cpu_baseline": {"value": value, "unit": "pairs/s", "cores": cores, "kind": "port",
                             "sample": "%d pairs per step at the full 1M-hypothesis budget, oracle/lr_oracle.c " % PAIRS_PER_STEP +
                                       "(OpenMP over hypotheses, fp64); the reference's native RANSAC "
                                       "(pygcransac / Open3D) is un-vendored and cannot be built here"},
            "e2e": {"value": value, "unit": "pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------- clocks
class ClockSampler(threading.Thread):
    """one streaming `nvidia-smi -lms 50` process; rows are parsed as they arrive"""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.rows, self.stop_flag, self.proc = index, [], False, None

    def run(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "50"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            for line in self.proc.stdout:
                parts = [x.strip() for x in line.strip().split(",")]
                if len(parts) >= 7:
                    self.rows.append(parts)
                if self.stop_flag:
                    break
        except Exception:
            pass

    def stop(self):
        self.stop_flag = True
        try:
            if self.proc:
                self.proc.terminate()
        except Exception:
            pass

    def summary(self):
        sm = [float(r[0]) for r in self.rows if r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for k, n in enumerate(names) if any(r[3 + k].lower().startswith("active") for r in self.rows)]
        busy = [v for v in sm if v > 0.5 * (max(mx) if mx else 1)]
        return {"sm_mhz": statistics.median(busy or sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(self.rows)}


# ----------------------------------------------------------------------------- our arm
def run_ours(args):
    import torch
    import torch.distributed as dist
    from lidarregistration_b200 import engine
    from lidarregistration_b200.algorithms import findRigidTransform

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("NCCL_DEBUG", "WARN")  # keeps NCCL's version banner off stdout (one JSON line)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)
    use_elc = not args.no_elc

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # pairs are sharded by rank (pair p -> rank p mod G); fixed work per GPU => weak scaling
    pairs = make_pairs(PAIRS_PER_STEP, CFG_SEED + rank * PAIRS_PER_STEP)
    host = [(torch.from_numpy(d["src"]).pin_memory(), torch.from_numpy(d["tgt"]).pin_memory()) for d in pairs]
    resident = [(a.to(dev), b.to(dev)) for a, b in host]
    params = engine.make_params(threshold=THRESH, confidence=1.0, max_iters=ITERS, seed=51, sample_size=3,
                                sampler=engine.SAMPLER_UNIFORM, use_elc=use_elc, refit=True)
    flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)

    def step_resident():  # the rank's pairs through the batched entry (two pairs in flight, one host sync)
        return engine.ransac_rigid_batch(resident, params)[-1]

    def step_e2e_batch():
        # the rank's pairs from pinned HOST memory through the C ABI's per-set entry (lr_ransac_rigid_batch = the loop of
        # Experiments/test.py:108-167 over the rank's pairs): the library copies pair i + 1 in under the kernels of pair i;
        # poses + statistics come back in pinned host memory, one synchronisation per step
        return engine.ransac_rigid_batch(host, params)[-1]

    def step_e2e():
        out = None
        for a, b in host:  # pinned HOST buffers through the reference-facing call (GC_RANSAC.py:46-49)
            out = findRigidTransform(a, b, threshold=THRESH, conf=1.0, spatial_coherence_weight=0.0,
                                     max_iters=ITERS, use_sprt=use_elc, min_inlier_ratio_for_sprt=-1,
                                     sampler=0, neighborhood=0, neighborhood_size=20, seed=51)
        return out

    def timed(fn, steps, warmup, prof=False):
        for _ in range(warmup):
            fn()
        barrier()
        if prof:
            engine.prof_read(engine.PROF_SCORE), engine.prof_read(engine.PROF_GEN)
            engine.prof_enable(True)
        tot = 0.0
        last = None
        for _ in range(steps):
            flush.fill_(1)
            barrier()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            last = fn()
            b.record()
            barrier()
            tot += a.elapsed_time(b)
        if prof:
            engine.prof_enable(False)
        return max_over_ranks(tot / steps), last

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ms_res, last = timed(step_resident, args.steps, args.warmup)

    # kernel accounting for the roofline: the same pairs one at a time (lr_ransac_rigid), so that the CUDA events
    # the library records around k_score on its stream see the kernel alone -- with two pairs in flight a queued
    # k_score would be charged the time it waits for the other pair's
    def step_single():
        out = None
        for a, b in resident:
            out = engine.ransac_rigid(a, b, params)
        return out

    prof_steps = max(2, args.steps // 2)
    ms_single, _ = timed(step_single, prof_steps, 1, prof=True)
    score_ms, score_launches = engine.prof_read(engine.PROF_SCORE)
    gen_ms, gen_launches = engine.prof_read(engine.PROF_GEN)
    ms_e2e_call, last_e2e = timed(step_e2e, args.steps, args.warmup)
    ms_e2e, last_e2e_b = timed(step_e2e_batch, args.steps, args.warmup)
    assert last_e2e_b["best_id"] == last["best_id"] and last_e2e_b["best_count"] == last["best_count"]
    sampler.stop()
    clocks = sampler.summary() if rank == 0 else None

    pairs_total = PAIRS_PER_STEP * world
    value = pairs_total / (ms_res * 1e-3)
    e2e_value = pairs_total / (ms_e2e * 1e-3)
    n_scored = last["n_scored"]
    # kernels launched by one lr_ransac_rigid call: k_ctl_reset + k_pack + per batch of <= 2^20 hypotheses (k_gen,
    # k_kabsch, k_score_tc, k_tc_events, k_resolve_end) + k_finish
    launches_per_pair = 2 + 5 * ((ITERS + (1 << 20) - 1) >> 20) + 1
    line = {"metric": METRIC, "value": value, "unit": "pairs/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_res, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32 (inlier sweep, bracketed) + f64 (models, recount)",
            "data": "synthetic", "config": config_dict(use_elc),
            "parallelism": "pairs sharded by rank (dp%d), no collective in the timed region" % world,
            "h_scored_per_pair": n_scored, "h_rechecked_per_pair": last["n_rechecked"],
            "best_count": last["best_count"],
            "e2e": {"value": e2e_value, "unit": "pairs/s", "ms_per_step": ms_e2e,
                    "h2d_bytes_per_step": PAIRS_PER_STEP * 2 * N_CORR * 12,
                    "d2h_bytes_per_step": PAIRS_PER_STEP * 600,
                    "api": "lr_ransac_rigid_batch (engine.ransac_rigid_batch): the step's pairs in pinned HOST memory in, "
                           "poses + refits + statistics in host memory out (one 600-byte control block per pair, written "
                           "by the last kernel); host->device copies by the copy engine inside the timed region, pair "
                           "i + 1 under the kernels of pair i; one synchronisation per step.  The reference's loop reads "
                           "only the pose of each pair (GC_RANSAC.py:46-55: the inlier mask is dropped)"},
            "e2e_per_pair_call": {"value": pairs_total / (ms_e2e_call * 1e-3), "unit": "pairs/s", "ms_per_step": ms_e2e_call,
                                  "h2d_bytes_per_step": PAIRS_PER_STEP * 2 * N_CORR * 12,
                                  "d2h_bytes_per_step": PAIRS_PER_STEP * (N_CORR + 2 * 128 + 120),
                                  "api": "findRigidTransform() once per pair (the signature of pygcransac.findRigidTransform, "
                                         "GC_RANSAC.py:46-49): pinned host arrays in, pose + inlier mask out, one "
                                         "synchronisation per pair -- rounds 1 and 2 reported this one as e2e"},
            "gpu_launches": launches_per_pair * PAIRS_PER_STEP * args.steps}

    # the same call with what the reference interface really hands over (GC_RANSAC.py:10-11): pageable numpy arrays
    host_np = [(d["src"], d["tgt"]) for d in pairs]

    def step_e2e_numpy():
        out = None
        for a, b in host_np:
            out = findRigidTransform(a, b, threshold=THRESH, conf=1.0, spatial_coherence_weight=0.0,
                                     max_iters=ITERS, use_sprt=use_elc, min_inlier_ratio_for_sprt=-1,
                                     sampler=0, neighborhood=0, neighborhood_size=20, seed=51)
        return out

    ms_e2e_np, _ = timed(step_e2e_numpy, max(2, args.steps // 2), 1)
    line["e2e"]["pinned_torch_inputs"] = True
    line["e2e_numpy"] = {"value": pairs_total / (ms_e2e_np * 1e-3), "unit": "pairs/s", "ms_per_step": ms_e2e_np,
                         "note": "same call, pageable numpy arrays in (staged by engine.to_dev_f32), numpy pose + mask out"}

    # second multi-GPU mode of the north star (BASELINE cfg 3 as worded): the hypotheses of ONE pair split across the
    # ranks; the round's packed (count, id) key is exchanged inside the kernel that ends the round, through peer
    # mailboxes over NVLink (lr_comm_init / lr_ransac_rigid_sharded); every rank regenerates the winner from the id
    hyp = bench_hypothesis_sharding(engine, torch, dist, resident, world, rank, barrier, max_over_ranks)
    if hyp is not None:
        line["hypothesis_sharding"] = hyp
        if args.mode == "hyp":
            h = hyp["elc_on"]
            line.update({"value": 1e3 / h["ms_per_pair"], "ms_per_step": h["ms_per_pair"], "scaling": "strong",
                         "steps": h["reps"], "warmup": 3,
                         "parallelism": "hypotheses of one pair sharded over %d ranks (peer-mailbox key exchange "
                                        "inside the round-end kernel)" % world})
            line["config"] = dict(line["config"], pairs_per_step=1)
    if rank == 0:
        line["clocks"] = clocks
        # ---- roofline of the dominant kernel: k_score_tc, the inlier sweep on tcgen05 (csrc/lr_score_tc.cuh)
        calls = prof_steps * PAIRS_PER_STEP
        h_per_launch = n_scored * calls / max(score_launches, 1)
        flops_per_launch = h_per_launch * N_CORR * FLOPS_PER_TEST          # 27 per residual test
        n_pad = (N_CORR + 127) // 128 * 128
        mma_flops_per_launch = ((h_per_launch + 127) // 128 * 128) * n_pad * 3 * 16 * 2  # three M128 x N x K16 MMAs per tile
        avg_ms = score_ms / max(score_launches, 1)
        achieved = flops_per_launch / (avg_ms * 1e-3) / 1e12 if avg_ms > 0 else 0.0
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        tensor_peak = peaks.get("bf16_tflops_sustained", 1369.0) if peaks else 1590.0
        fp32_peak = engine.peak_fp32(0)
        sm_hz = (clocks["sm_mhz"] or 1965.0) * 1e6
        # the epilogue is what the sweep is made of: 3.25 issue slots per residual (3 packed FMAs per two residuals,
        # one sign-bit add, half a 3-input min), 32 residuals per warp instruction, 4 schedulers per SM
        issue_frac = (h_per_launch * N_CORR * 3.25 / 32) / (148 * 4 * sm_hz * avg_ms * 1e-3) if avg_ms > 0 else None
        line["roofline"] = {
            "kernel": "k_score_tc (inlier sweep: tcgen05 residual components, TMEM -> register epilogue, banded fp64 recheck)"
                      " + k_tc_events",
            "bound": "tensor",
            "achieved": achieved, "peak": tensor_peak, "unit": "TFLOP/s", "frac": achieved / tensor_peak,
            "peak_source": ("bf16_tflops_sustained of MEASURED_PEAKS.json (kernel timed inside a step)" if peaks else
                            "fallback of B200_PROFILING.md"),
            "note": "achieved counts the ALGORITHMIC 27 FLOP per residual test (SURVEY 8(d)); the MMAs execute 96 "
                    "(K padded to 16, three fp16 pieces) -> mma_tflops.  Neither figure is the limiter: the sweep is bound "
                    "by the issue slots of its TMEM -> register epilogue (issue_frac) and by the TMEM round trip "
                    "(DESIGN.md 5.1); fp32_equiv_frac compares with the CUDA-core FP32 roofline round 1 was bound by",
            "mma_tflops": mma_flops_per_launch / (avg_ms * 1e-3) / 1e12 if avg_ms > 0 else None,
            "mma_frac": mma_flops_per_launch / (avg_ms * 1e-3) / 1e12 / tensor_peak if avg_ms > 0 else None,
            "issue_frac": issue_frac,
            "fp32_equiv_frac": achieved / fp32_peak if fp32_peak else None, "fp32_peak_probe": fp32_peak,
            "traffic": ncu_traffic("r2_ncu_k_score_tc.txt"),
            "traffic_note": "DRAM bytes of one ncu --set full launch (profiles/r2_ncu_k_score_tc.txt); algorithmic bytes per "
                            "launch = n x 64 B of correspondence operands + H_scored x (96 B operand image + 96 B fp64 model "
                            "+ 8 B band / count), every CTA re-reads the correspondences from L2",
            "avg_launch_ms": avg_ms, "launches": score_launches,
            "share_of_step": score_ms / (ms_single * prof_steps) if ms_single else None,
            "gen_share_of_step": gen_ms / (ms_single * prof_steps) if ms_single else None,
            "share_note": "shares and avg_launch_ms from a pass over the same pairs one at a time (%.3f ms per step); "
                          "`value` runs them through lr_ransac_rigid_batch (two pairs in flight)" % ms_single,
            "algorithmic_flops_per_launch": flops_per_launch}
        if not args.skip_extras:
            line["mnn_match"] = bench_matching(engine, torch, dev)
            other = bench_other_regime(engine, torch, resident, not use_elc)
            line["other_regime"] = other
            try:  # torchrun exports OMP_NUM_THREADS=1: the CPU legs use every core this rank may run on
                from oracle import lr_oracle as _O
                _O.set_threads(len(os.sched_getaffinity(0)))
            except Exception:
                pass
            line["gc_semantics"] = bench_gc_semantics(engine, torch, resident, pairs[0], use_elc)
            cpu_pairs = make_pairs(1, CFG_SEED) * 8
            cpu_ransac_rate(cpu_pairs[:1], use_elc)  # warm the thread pool
            v, cores, secs = cpu_ransac_rate(cpu_pairs, use_elc)
            line["cpu_baseline"] = {"value": v, "unit": "pairs/s", "cores": cores, "kind": "port",
                                    "sample": "8 passes over one cfg-3 pair at the full 1M-hypothesis budget (%.1f s wall, "
                                              "%.0f core-seconds), oracle/lr_oracle.c, OpenMP over hypotheses, fp64"
                                              % (secs, secs * cores)}
            line["speedup_vs_cpu"] = {"e2e": e2e_value / world / v, "resident": value / world / v, "cores": cores,
                                      "note": "per GPU, this run's own cpu_baseline; the >= 100x target is host dependent "
                                              "(round 1: 143x on a 16-core box, 82x on a 32-core box)"}
            line["reference_faithful"] = bench_faithful_regime(engine, torch, resident, pairs[0])
            line["fr_e2e"] = bench_fr(torch)
            line["f4_consumers"] = bench_f4(engine, torch, pairs[0])
            line["reference_libraries"] = probe_reference_libraries(pairs[0])
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def bench_hypothesis_sharding(engine, torch, dist, resident, world, rank, barrier, max_over_ranks):
    """one cfg-3 pair, 1M hypotheses split over the ranks; ELC on (the headline regime) and off (every hypothesis
    scored).  Rank 0 also runs the same pair through the 1-GPU entry: the N-rank result must be identical
    (best id, best count, T, T_refit bit for bit) and the ratio of the two times is the strong-scaling speed-up."""
    if world < 2:
        return None
    from lidarregistration_b200 import parallel
    a0, b0 = (t.clone() for t in resident[0])
    for t in (a0, b0):
        dist.broadcast(t, src=0)
    p2p = parallel.init_comm()
    out = {"transport": "p2p (cudaIpc peer mailboxes, exchange fused into k_resolve_end)" if p2p else
           "allreduce (torch.distributed MAX over NCCL)", "world": world}
    for name, elc in (("elc_on", True), ("elc_off", False)):
        params = engine.make_params(threshold=THRESH, confidence=1.0, max_iters=ITERS, seed=51, sample_size=3,
                                    sampler=engine.SAMPLER_UNIFORM, use_elc=elc, refit=True)
        reps = 20 if elc else 5
        for _ in range(3):
            parallel.ransac_rigid_sharded(a0, b0, params)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            res = parallel.ransac_rigid_sharded(a0, b0, params)
        e1.record()
        barrier()
        hyp_ms = max_over_ranks(e0.elapsed_time(e1) / reps)
        # the 1-GPU entry on the same pair, timed on rank 0 while the other ranks wait at the barrier
        single_ms, same = 0.0, True
        if rank == 0:
            for _ in range(3):
                ref = engine.ransac_rigid(a0, b0, params)
            torch.cuda.synchronize()
            e0.record()
            for _ in range(reps):
                ref = engine.ransac_rigid(a0, b0, params)
            e1.record()
            torch.cuda.synchronize()
            single_ms = e0.elapsed_time(e1) / reps
            same = bool(res["best_id"] == ref["best_id"] and res["best_count"] == ref["best_count"] and
                        res["iters_run"] == ref["iters_run"] and (res["T"] == ref["T"]).all() and
                        (res["T_refit"] == ref["T_refit"]).all())
        barrier()
        single_ms = max_over_ranks(single_ms)
        ok = torch.tensor([1 if same else 0], device=a0.device)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        assert int(ok.item()) == 1, "hypothesis sharding over %d ranks differs from the 1-GPU result (%s)" % (world, name)
        out[name] = {"ms_per_pair": hyp_ms, "pairs_per_s": 1e3 / hyp_ms, "ms_per_pair_1gpu": single_ms,
                     "speedup": single_ms / hyp_ms, "efficiency": single_ms / hyp_ms / world, "reps": reps,
                     "identical_to_1gpu": True, "best_count": res["best_count"], "best_id": res["best_id"],
                     "h_scored_all_ranks": res["n_scored"]}
    return out


def bench_faithful_regime(engine, torch, resident, pair0):
    """SURVEY 8(d) cfg 3, "reference-faithful": the confidence exit on (0.9995) on both sides, and the Open3D branch's
    semantics (FR.py:122-139: ransac_n = 4 drawn with replacement, edge-length checker) timed on the CPU in its two
    forms -- *faithful* (copy and transform the whole source cloud per surviving hypothesis, as upstream) and *lean*
    (transform the correspondences only)."""
    from oracle import lr_oracle as O
    a, b = resident[0]
    out = {}
    for name, kw in (("gc_count_m3", dict(sample_size=3, sampler=engine.SAMPLER_UNIFORM)),
                     ("open3d_m4_replace", dict(sample_size=4, sampler=engine.SAMPLER_REPLACE))):
        params = engine.make_params(threshold=THRESH, confidence=0.9995, max_iters=ITERS, seed=51, use_elc=True,
                                    refit=True, **kw)
        for _ in range(3):
            res = engine.ransac_rigid(a, b, params)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        reps = 20
        e0.record()
        for _ in range(reps):
            res = engine.ransac_rigid(a, b, params)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        row = {"gpu_ms_per_pair": ms, "gpu_pairs_per_s": 1e3 / ms, "iters_run": res["iters_run"],
               "best_count": res["best_count"], "best_id": res["best_id"]}
        okw = dict(m=kw["sample_size"], sampler=0 if kw["sampler"] == engine.SAMPLER_UNIFORM else 2, use_elc=True,
                   thr=THRESH, conf=0.9995, max_iters=ITERS, round_size=65536, seed=51, refit=True)
        for variant, faithful in (("cpu_lean", False), ("cpu_faithful", True)):
            if faithful and kw["sample_size"] == 3:
                continue
            O.set_o3d_faithful(faithful)
            try:
                O.ransac(pair0["src"], pair0["tgt"], **okw)  # warm
                t0 = time.perf_counter()
                n = 5
                for _ in range(n):
                    o = O.ransac(pair0["src"], pair0["tgt"], **okw)
                dt = (time.perf_counter() - t0) / n
            finally:
                O.set_o3d_faithful(False)
            row[variant] = {"ms_per_pair": dt * 1e3, "pairs_per_s": 1.0 / dt, "cores": O.num_threads(),
                            "same_selection": bool(o["best_id"] == res["best_id"] and o["best_count"] == res["best_count"]
                                                   and o["iters_run"] == res["iters_run"])}
        out[name] = row
    # the fixed-budget Open3D-semantics CPU variants (the >= 100x denominator of SURVEY 8(d) is the faster of the two)
    okw = dict(m=4, sampler=2, use_elc=True, thr=THRESH, conf=1.0, max_iters=ITERS, round_size=65536, seed=51, refit=True)
    fixed = {}
    for variant, faithful in (("cpu_lean", False), ("cpu_faithful", True)):
        O.set_o3d_faithful(faithful)
        try:
            t0 = time.perf_counter()
            O.ransac(pair0["src"], pair0["tgt"], **okw)
            dt = time.perf_counter() - t0
        finally:
            O.set_o3d_faithful(False)
        fixed[variant] = {"ms_per_pair": dt * 1e3, "pairs_per_s": 1.0 / dt, "cores": O.num_threads()}
    params = engine.make_params(threshold=THRESH, confidence=1.0, max_iters=ITERS, seed=51, use_elc=True, refit=True,
                                sample_size=4, sampler=engine.SAMPLER_REPLACE)
    for _ in range(3):
        res = engine.ransac_rigid(a, b, params)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(10):
        res = engine.ransac_rigid(a, b, params)
    e1.record()
    torch.cuda.synchronize()
    fixed["gpu_ms_per_pair"] = e0.elapsed_time(e1) / 10
    fixed["gpu_h_scored"] = res["n_scored"]
    fixed["gpu_over_faster_cpu_variant"] = min(fixed["cpu_lean"]["ms_per_pair"], fixed["cpu_faithful"]["ms_per_pair"]) / \
        fixed["gpu_ms_per_pair"]
    out["open3d_m4_replace_fixed_budget"] = fixed
    return out


def bench_f4(engine, torch, pair0):
    """SURVEY 8(f4), the two consumers right after the path: the ICP refinement of Experiments/test.py:183-188 on a
    25k-point pair (device-resident: hashed-grid nearest neighbour, one kernel per iteration) and PointDSC's seed scoring
    (Experiments/models/PointDSC.py:319-336) on the cfg-3 pair with 3 000 seed transforms (ratio 0.1 of 30 000), each
    beside the oracle on the host cores."""
    import numpy as np
    from lidarregistration_b200 import synthetic
    from lidarregistration_b200.algorithms import registration_icp, registration_icp_bruteforce
    from oracle import lr_oracle as O
    out = {}
    p = synthetic.make_pair(25000, seed=51 + 5000, overlap=0.6)
    ang = np.deg2rad(1.0)
    d = np.eye(4)
    d[:2, :2] = [[np.cos(ang), -np.sin(ang)], [np.sin(ang), np.cos(ang)]]
    d[:3, 3] = [0.2, -0.15, 0.05]
    T0 = d @ p["T_gt"]
    a, b = engine.to_dev_f32(p["xyz0"]), engine.to_dev_f32(p["xyz1"])

    def wall(fn, reps):
        fn()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(reps):
            r = fn()
        torch.cuda.synchronize()
        return (time.perf_counter() - t0) / reps * 1e3, r

    ms_grid, r = wall(lambda: registration_icp(a, b, 0.6, T0), 20)
    ms_bf, rb = wall(lambda: registration_icp_bruteforce(a, b, 0.6, T0), 2)
    t0 = time.perf_counter()
    To, fo, ro, ito = O.icp(p["xyz0"], p["xyz1"], 0.6, T0, max_iteration=2)
    cpu_eval_ms = (time.perf_counter() - t0) / 3 * 1e3
    out["icp"] = {"workload": "registration_icp(src, tgt, 0.6, T_init): 25k-point pair, T_init = T_gt perturbed by 1 deg / 0.25 m",
                  "ms": ms_grid, "iterations": r.iterations, "fitness": r.fitness, "inlier_rmse": r.inlier_rmse,
                  "ms_per_iteration": ms_grid / (r.iterations + 1),
                  "bruteforce_sweep_ms": ms_bf, "bruteforce_iterations": rb.iterations,
                  "cpu_oracle_ms_per_iteration": cpu_eval_ms, "cpu_cores": O.num_threads(),
                  "cpu_note": "oracle/lr_oracle.c brute-force radius search + Kabsch, 3 evaluations timed"}
    src, tgt = pair0["src"], pair0["tgt"]
    rng = np.random.default_rng(7)
    S = 3000
    models = np.tile(pair0["T_gt"], (S, 1, 1))
    models[:, :3, 3] += rng.normal(0, 0.5, (S, 3)) * (rng.random(S) < 0.7)[:, None]
    ds, dt = engine.to_dev_f32(src), engine.to_dev_f32(tgt)
    dm = torch.from_numpy(models).to(ds.device)
    ms_seed, res = wall(lambda: engine.seeds_score(ds, dt, dm, THRESH), 20)
    t0 = time.perf_counter()
    counts, best = O.seeds_score(src, tgt, models, THRESH)
    cpu_ms = (time.perf_counter() - t0) * 1e3
    assert res["best"] == best and int(res["counts"][best]) == int(counts[best])
    out["seed_scoring"] = {"workload": "3 000 seed transforms x 30 000 correspondences (cfg-3 pair), inlier threshold 0.6 m",
                           "ms": ms_seed, "best": res["best"], "best_count": res["best_count"],
                           "cpu_oracle_ms": cpu_ms, "cpu_cores": O.num_threads(), "identical_to_oracle": True}
    return out


def bench_fr(torch):
    """The drop-in entry itself (FR.py:16-119): one 25k-point synthetic pair, --mode MMN --iters 1000000, host tensors in,
    T out; stage breakdown from CUDA events around the same calls FR() makes."""
    from types import SimpleNamespace
    from lidarregistration_b200 import engine, synthetic
    from lidarregistration_b200.algorithms import FR
    p = synthetic.make_pair(25000, seed=51 + 5000, overlap=0.6)
    t = [torch.from_numpy(p[k]) for k in ("xyz0", "xyz1", "feat0", "feat1")]
    args = SimpleNamespace(mode="MMN", iters=ITERS, codebase="GC", prosac=True, spatial_coherence_weight=0.0, GC_conf=1.0,
                           fast_rejection="ELC", GC_LO=True, GPF_factor=2.0, GPF_grid_wid=10, GPF_max_matches=10 ** 9, seed=51)
    for _ in range(3):
        out = FR(*t, args, p["T_gt"])
    torch.cuda.synchronize()
    reps = 10
    t0 = time.perf_counter()
    model_time = 0.0
    for _ in range(reps):
        out = FR(*t, args, p["T_gt"])
        model_time += out[1]
    torch.cuda.synchronize()
    wall = (time.perf_counter() - t0) / reps
    # stages on device tensors
    f0, f1 = engine.to_dev_f32(t[2]), engine.to_dev_f32(t[3])
    x0, x1 = engine.to_dev_f32(t[0]), engine.to_dev_f32(t[1])
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(5)]
    stage = [0.0, 0.0, 0.0, 0.0]
    for _ in range(reps):
        ev[0].record()
        i1, i2 = engine.match_nn(f0, f1, want_2nd=True)
        ev[1].record()
        mi, mj = engine.match_mutual(f0, f1, i1)
        ev[2].record()
        src, tgt = engine.gather_xyz(x0, mi), engine.gather_xyz(x1, mj)
        q = engine.match_ratio(f0, f1, mi, mj, i2[mi])
        order = torch.argsort(q, stable=True)
        src, tgt = src[order].contiguous(), tgt[order].contiguous()
        ev[3].record()
        params = engine.make_params(threshold=THRESH, confidence=1.0, max_iters=ITERS, seed=51, sample_size=3,
                                    sampler=engine.SAMPLER_PROSAC, use_elc=True, refit=True)
        res = engine.ransac_rigid(src, tgt, params)
        ev[4].record()
        torch.cuda.synchronize()
        for k in range(4):
            stage[k] += ev[k].elapsed_time(ev[k + 1]) / reps
    from lidarregistration_b200 import metrics
    return {"workload": "FR(A, B, A_feat, B_feat, args, T_gt): 25k-point pair, --mode MMN --iters 1000000 --prosac True, "
                        "host torch tensors in, numpy T out", "wall_ms_per_pair": wall * 1e3,
            "model_time_ms": model_time / reps * 1e3, "pairs_per_s": 1.0 / wall,
            "stages_ms": {"find_2nn (one sweep, both neighbours)": stage[0], "nn_to_mutual": stage[1],
                          "gather + ratio quality + PROSAC sort": stage[2], "RANSAC + refit": stage[3]},
            "mutual_pairs": int(mi.shape[0]), "h_scored": res["n_scored"],
            "registration_success": bool(metrics.registration_success(out[0], p["T_gt"]))}


def probe_reference_libraries(pair0):
    """SURVEY 8(c), last row: when the reference's third-party engines are importable on the box, time the real thing."""
    out = {}
    for mod in ("open3d", "pygcransac"):
        try:
            __import__(mod)
            out[mod] = "importable"
        except Exception as e:
            out[mod] = "absent (%s)" % type(e).__name__
    if out.get("pygcransac") == "importable":
        try:
            import numpy as np
            import pygcransac
            t0 = time.perf_counter()
            pose, mask = pygcransac.findRigidTransform(np.ascontiguousarray(pair0["src"]), np.ascontiguousarray(pair0["tgt"]),
                                                       threshold=THRESH, conf=0.9995, spatial_coherence_weight=0.0,
                                                       max_iters=ITERS, use_sprt=True, min_inlier_ratio_for_sprt=-1,
                                                       sampler=0, neighborhood=0, neighborhood_size=20)
            out["pygcransac_ms_per_pair"] = (time.perf_counter() - t0) * 1e3
            out["pygcransac_inliers"] = int(mask.sum())
        except Exception as e:
            out["pygcransac_error"] = repr(e)
    if out.get("open3d") == "importable":
        try:
            import numpy as np
            import open3d as o3d
            pc0, pc1 = o3d.geometry.PointCloud(), o3d.geometry.PointCloud()
            pc0.points = o3d.utility.Vector3dVector(pair0["src"].astype(np.float64))
            pc1.points = o3d.utility.Vector3dVector(pair0["tgt"].astype(np.float64))
            n = len(pair0["src"])
            corres = o3d.utility.Vector2iVector(np.stack([np.arange(n), np.arange(n)], 1).astype(np.int32))
            reg = o3d.pipelines.registration
            t0 = time.perf_counter()
            r = reg.registration_ransac_based_on_correspondence(
                pc0, pc1, corres, THRESH, reg.TransformationEstimationPointToPoint(False), 4,
                [reg.CorrespondenceCheckerBasedOnEdgeLength(0.9)], reg.RANSACConvergenceCriteria(ITERS, 0.9995))
            out["open3d_ms_per_pair"] = (time.perf_counter() - t0) * 1e3
            out["open3d_fitness"] = float(r.fitness)
        except Exception as e:
            out["open3d_error"] = repr(e)
    return out


def ncu_traffic(summary):
    """dram__bytes_read.sum + dram__bytes_write.sum (bytes per launch) from a committed ncu summary, else None"""
    try:
        tot, unit = 0.0, {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
        for line in open(os.path.join(ROOT, "profiles", summary)):
            f = line.split()
            if f and f[0] in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
                tot += float(f[1]) * unit.get(f[2], 1.0)
        return tot or None
    except Exception:
        return None


def bench_other_regime(engine, torch, resident, use_elc):
    """the regime that is not the headline (ELC off = every hypothesis scored = the FP32 roofline case)"""
    params = engine.make_params(threshold=THRESH, confidence=1.0, max_iters=ITERS, seed=51, use_elc=use_elc)
    a, b = resident[0]
    for _ in range(2):
        engine.ransac_rigid(a, b, params)
    engine.prof_read(engine.PROF_SCORE)
    engine.prof_enable(True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    reps = 3
    for _ in range(reps):
        res = engine.ransac_rigid(a, b, params)
    e1.record()
    torch.cuda.synchronize()
    engine.prof_enable(False)
    ms = e0.elapsed_time(e1) / reps
    sms, sl = engine.prof_read(engine.PROF_SCORE)
    flops = res["n_scored"] * reps * N_CORR * FLOPS_PER_TEST
    return {"elc": use_elc, "pairs_per_s": 1e3 / ms, "ms_per_pair": ms, "h_scored_per_pair": res["n_scored"],
            "k_score_tflops": flops / (sms * 1e-3) / 1e12 if sms > 0 else None, "k_score_ms_per_pair": sms / reps}


def bench_gc_semantics(engine, torch, resident, pair0, use_elc):
    """the same cfg-3 pair under pygcransac's own criterion -- quantised MSAC selection in fp64
    (k_score_msac), 10 x 20 local-optimisation draws, 10 passes of iterated least squares -- with the oracle's
    lro_ransac_gc timed beside it on one pass."""
    params = engine.make_params(threshold=THRESH, confidence=1.0, max_iters=ITERS, seed=51, use_elc=use_elc,
                                scoring=engine.SCORE_MSAC, lo_rounds=10, lo_trials=20, lsq_iters=10)
    a, b = resident[0]
    for _ in range(2):
        engine.ransac_rigid(a, b, params)
    engine.prof_read(engine.PROF_SCORE)
    engine.prof_enable(True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    reps = 5
    for _ in range(reps):
        res = engine.ransac_rigid(a, b, params)
    e1.record()
    torch.cuda.synchronize()
    engine.prof_enable(False)
    ms = e0.elapsed_time(e1) / reps
    sms, _ = engine.prof_read(engine.PROF_SCORE)
    out = {"scoring": "MSAC at 1.5 x threshold, integer-quantised (include/lidarreg.h LR_SCORE_MSAC)",
           "pairs_per_s": 1e3 / ms, "ms_per_pair": ms, "k_score_msac_ms_per_pair": sms / reps,
           "h_scored_per_pair": res["n_scored"],
           "stats": {k: res[k] for k in ("best_id", "best_count", "best_score", "lo_score", "final_score",
                                         "lo_improved", "lsq_improved")}}
    try:
        from oracle import lr_oracle as O
        t0 = time.perf_counter()
        o = O.ransac_gc(pair0["src"], pair0["tgt"], thr=THRESH, conf=1.0, max_iters=ITERS, seed=51, use_elc=use_elc)
        out["cpu_baseline"] = {"ms_per_pair": (time.perf_counter() - t0) * 1e3, "cores": O.num_threads(),
                               "kind": "port", "same_selection": bool(o["best_id"] == res["best_id"] and
                                                                      o["best_score"] == res["best_score"] and
                                                                      o["lo_score"] == res["lo_score"])}
    except Exception as e:  # the checker is optional here, the measurement is not
        out["cpu_baseline"] = {"error": repr(e)}
    return out


def bench_matching(engine, torch, dev):
    """cfg 2: mutual-NN matching, N = M = 50k x 32 (matching.py:22-65 + :222-239), device resident."""
    g = torch.Generator(device=dev).manual_seed(51 + 2000)
    f0 = torch.nn.functional.normalize(torch.randn(MATCH_N, 32, device=dev, generator=g), dim=1)
    f1 = torch.nn.functional.normalize(torch.randn(MATCH_N, 32, device=dev, generator=g), dim=1)
    k = MATCH_N // 2
    f1[:k] = torch.nn.functional.normalize(f0[:k] + 0.08 * torch.randn(k, 32, device=dev, generator=g), dim=1)

    def run():
        i1, _ = engine.match_nn(f0, f1, want_2nd=False)
        return engine.match_mutual(f0, f1, i1)

    for _ in range(3):
        mi, _ = run()
    engine.prof_read(engine.PROF_NN)
    engine.prof_enable(True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    reps = 5
    for _ in range(reps):
        run()
    e1.record()
    torch.cuda.synchronize()
    engine.prof_enable(False)
    ms = e0.elapsed_time(e1) / reps
    nn_ms, nn_l = engine.prof_read(engine.PROF_NN)
    # find_nn alone (operand image + sweep + exact re-rank), one direction
    torch.cuda.synchronize()
    e0.record()
    for _ in range(reps):
        engine.match_nn(f0, f1, want_2nd=False)
    e1.record()
    torch.cuda.synchronize()
    find_nn_ms = e0.elapsed_time(e1) / reps
    flops = 2.0 * MATCH_N * MATCH_N * 32  # per sweep; MNN = forward + reverse sweep
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    tensor_peak = peaks.get("bf16_tflops", 1590.0)
    ach = flops / (nn_ms / max(nn_l, 1) * 1e-3) / 1e12 if nn_ms > 0 else None
    # CPU baseline of the same sweep on a bounded sample: 4096 query rows against all 50k targets
    from oracle import lr_oracle as O
    rows = 4096
    h0, h1 = f0[:rows].cpu().numpy(), f1.cpu().numpy()
    t0 = time.perf_counter()
    O.find_nn(h0, h1)
    cpu_s = time.perf_counter() - t0
    cpu_mnn_ms = cpu_s * (MATCH_N / rows) * 2 * 1e3
    return {"ms": ms, "cpu_baseline": {"ms": cpu_mnn_ms, "cores": O.num_threads(), "kind": "port",
                                       "sample": "oracle find_nn, %d of %d query rows x all targets (%.2f s), scaled to "
                                                 "two full sweeps" % (rows, MATCH_N, cpu_s)}, "unit": "ms per mutual-NN match (forward + reverse sweep + intersection), N=M=50000, D=32",
            "mutual_pairs": int(mi.shape[0]), "sweep_ms": nn_ms / max(nn_l, 1), "find_nn_ms": find_nn_ms,
            "fractions_of_tensor_peak": {"sweep (k_nn_tc alone)": ach / tensor_peak if ach else None,
                                         "find_nn (prep + sweep + re-rank)": flops / (find_nn_ms * 1e-3) / 1e12 / tensor_peak,
                                         "mutual match (two sweeps of 2 N M D + intersection)": 2 * flops / (ms * 1e-3) / 1e12 / tensor_peak},
            "roofline": {"kernel": "k_nn_tc alone (tcgen05 sweep: fp16 operands, fp16 accumulators in TMEM, chunk-maximum events); the library brackets this kernel only", "bound": "tensor", "achieved": ach,
                         "peak": tensor_peak, "unit": "TFLOP/s", "frac": ach / tensor_peak if ach else None,
                         "peak_source": "bf16_tflops of MEASURED_PEAKS.json (burst)" if peaks else "fallback 1590",
                         "traffic": ncu_traffic("r1_ncu_k_nn_tc.txt")}}


def main():
    args = parse()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
