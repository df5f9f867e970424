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
PAIRS_PER_STEP = 4
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
    return ap.parse_args()


def config_dict(use_elc):
    return {"workload": "cfg3: RANSAC --iters 1000000 on one 30k-correspondence synthetic pair, 30%% inliers, "
                        "3-point samples, ELC %s, fixed budget (confidence exit off), count scoring + LSQ refit"
                        % ("on (reference default --fast_rejection ELC)" if use_elc else "off (every hypothesis scored)"),
            "n_correspondences": N_CORR, "iters": ITERS, "inlier_ratio": INLIER_RATIO, "threshold_m": THRESH,
            "pairs_per_step": PAIRS_PER_STEP, "elc": bool(use_elc),
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
    pairs = make_pairs(1, CFG_SEED)
    from oracle import lr_oracle as O
    # all the host threads this process may use (torchrun exports OMP_NUM_THREADS=1 for its workers)
    try:
        O.set_threads(len(os.sched_getaffinity(0)))
    except (AttributeError, OSError):
        O.set_threads(os.cpu_count() or 1)
    cores = O.num_threads()
    for _ in range(min(args.warmup, 1)):
        cpu_ransac_rate(pairs, use_elc)
    times = []
    for _ in range(args.steps):
        _, _, dt = cpu_ransac_rate(pairs, use_elc)
        times.append(dt)
    ms = 1e3 * sum(times) / len(times)
    value = 1e3 / ms
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": "pairs/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": min(args.warmup, 1), "ms_per_step": ms, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": config_dict(use_elc),
            "cpu_baseline": {"value": value, "unit": "pairs/s", "cores": cores, "kind": "port",
                             "sample": "1 pair per step at the full 1M-hypothesis budget, oracle/lr_oracle.c "
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
    ms_e2e, last_e2e = timed(step_e2e, args.steps, args.warmup)
    sampler.stop()
    clocks = sampler.summary() if rank == 0 else None

    pairs_total = PAIRS_PER_STEP * world
    value = pairs_total / (ms_res * 1e-3)
    e2e_value = pairs_total / (ms_e2e * 1e-3)
    n_scored = last["n_scored"]
    # kernels launched by one lr_ransac_rigid call: reset + pack + per batch (gen, score, resolve, recount,
    # round_end) + model_from_key + mask_sums + refit_H + refit_solve
    launches_per_pair = 2 + 5 * ((ITERS + (1 << 20) - 1) >> 20) + 4
    line = {"metric": METRIC, "value": value, "unit": "pairs/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_res, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32 (inlier sweep, bracketed) + f64 (models, recount)",
            "data": "synthetic", "config": dict(config_dict(use_elc), parallelism="pairs sharded by rank (dp%d), "
                                                "no collective in the timed region" % world),
            "h_scored_per_pair": n_scored, "h_rechecked_per_pair": last["n_rechecked"],
            "best_count": last["best_count"],
            "e2e": {"value": e2e_value, "unit": "pairs/s", "ms_per_step": ms_e2e,
                    "h2d_bytes_per_step": PAIRS_PER_STEP * 2 * N_CORR * 12,
                    "d2h_bytes_per_step": PAIRS_PER_STEP * (N_CORR + 2 * 128 + 120)},
            "gpu_launches": launches_per_pair * PAIRS_PER_STEP * args.steps}

    # second multi-GPU mode of the north star: the hypotheses of ONE pair split across the ranks,
    # one 8-byte NCCL MAX all-reduce of the packed (count, id) key per round
    if world > 1:
        from lidarregistration_b200 import parallel
        a0, b0 = (t.clone() for t in resident[0])
        for t in (a0, b0):
            dist.broadcast(t, src=0)
        for _ in range(3):
            parallel.ransac_rigid_sharded(a0, b0, params)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 10
        e0.record()
        for _ in range(reps):
            shard_res = parallel.ransac_rigid_sharded(a0, b0, params)
        e1.record()
        barrier()
        hyp_ms = max_over_ranks(e0.elapsed_time(e1) / reps)
        line["hypothesis_sharding"] = {"ms_per_pair": hyp_ms, "pairs_per_s": 1e3 / hyp_ms, "best_count": shard_res["best_count"],
                                       "note": "one pair, 1M hypotheses split over %d ranks, 1 all-reduce(MAX, 8 B)" % world}
    if rank == 0:
        line["clocks"] = clocks
        # ---- roofline of the dominant kernel (k_score, the inlier sweep): SM FP32-bound
        calls = prof_steps * PAIRS_PER_STEP
        flops_per_launch = (n_scored * calls / max(score_launches, 1)) * N_CORR * FLOPS_PER_TEST
        avg_ms = score_ms / max(score_launches, 1)
        achieved = flops_per_launch / (avg_ms * 1e-3) / 1e12 if avg_ms > 0 else 0.0
        peak = engine.peak_fp32(0)
        peak2 = engine.peak_fp32(1)
        line["roofline"] = {
            "kernel": "k_score (inlier sweep)", "bound": "fp32",
            "note": "SM FP32-FMA bound (SURVEY 8(d)); neither HBM nor tensor: 24 B/correspondence live in "
                    "shared memory and are reused by every hypothesis",
            "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak if peak else None,
            "peak_source": "FFMA probe measured in this run (lr_peak_fp32 mode 0); packed fma.rn.f32x2 probe: "
                           "%.1f TFLOP/s; nominal 148 SM x 128 lanes x 2 x %.0f MHz = %.1f" %
                           (peak2, clocks["sm_max_mhz"] or 1965, 148 * 128 * 2 * (clocks["sm_max_mhz"] or 1965) / 1e6),
            "traffic": ncu_traffic("r1_ncu_k_score.txt"),
            "traffic_note": "DRAM bytes of one ncu --set full launch (profiles/r1_ncu_k_score.txt); the kernel is not "
                            "memory bound: algorithmic bytes per launch = n x 48 B of correspondences + H_scored x 64 B of "
                            "models",
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
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


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
            "mutual_pairs": int(mi.shape[0]), "sweep_ms": nn_ms / max(nn_l, 1),
            "roofline": {"kernel": "k_nn_tc (tcgen05 sweep: fp16 operands, fp16 accumulators in TMEM, chunk-maximum events) + k_rerank (exact fp32)", "bound": "tensor", "achieved": ach,
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
