"""A/B of the matching sweep implementations on the cfg-2 workload (50k x 50k x 32):
mode 1 = exact CUDA-core sweep (the on-device truth, itself pinned to the oracle by the tests),
mode 2 / 3 = tcgen05 sweep with fp32 / fp16 accumulators.  Prints per-sweep kernel time and index parity.
LR_SO=<path> loads another build of the library (e.g. one compiled with -DLR_TC_TIMING)."""
import os
import sys
import ctypes

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from lidarregistration_b200 import engine, _lib, build  # noqa: E402

if os.environ.get("LR_SO"):
    build.SO = os.environ["LR_SO"]


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 50000
    dev = torch.device("cuda:0")
    g = torch.Generator(device=dev).manual_seed(51 + 2000)
    f0 = torch.nn.functional.normalize(torch.randn(n, 32, device=dev, generator=g), dim=1)
    f1 = torch.nn.functional.normalize(torch.randn(n, 32, device=dev, generator=g), dim=1)
    k = n // 2
    f1[:k] = torch.nn.functional.normalize(f0[:k] + 0.08 * torch.randn(k, 32, device=dev, generator=g), dim=1)
    truth = None
    L = _lib.lib()
    for mode in ((2, 3) if os.environ.get("LR_SO") else (1, 2, 3)):
        engine.match_set_mode(mode)
        for want2 in (False, True):
            for _ in range(3):
                i1, i2 = engine.match_nn(f0, f1, want_2nd=want2)
            engine.prof_read(engine.PROF_NN)
            engine.prof_enable(True)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            if hasattr(L, "lr_tc_timing_dump"):
                L.lr_tc_timing_dump()  # reset
            e0.record()
            reps = 10
            for _ in range(reps):
                i1, i2 = engine.match_nn(f0, f1, want_2nd=want2)
            e1.record()
            torch.cuda.synchronize()
            engine.prof_enable(False)
            ms, cnt = engine.prof_read(engine.PROF_NN)
            if truth is None:
                truth = {}
            key = want2
            if mode == 1:
                truth[key] = (i1.clone(), None if i2 is None else i2.clone())
                ok = "truth"
            elif key not in truth:
                ok = "(no truth in this run)"
            else:
                t1, t2 = truth[key]
                ok = "idx1 %s" % bool(torch.equal(i1, t1)) + ("" if i2 is None else ", idx2 %s" % bool(torch.equal(i2, t2)))
            sweep = ms / max(cnt, 1)
            print("mode %d want2 %d: sweep kernel %.4f ms (%.0f TFLOP/s algorithmic), whole match_nn %.4f ms | %s" %
                  (mode, want2, sweep, 2.0 * n * n * 32 / (sweep * 1e-3) / 1e12, e0.elapsed_time(e1) / reps, ok), flush=True)
            if mode != 1 and hasattr(L, "lr_tc_trace_dump") and not want2:
                sys.stderr.flush()
                L.lr_tc_trace_dump()
            if mode != 1 and hasattr(L, "lr_tc_timing_dump"):
                sys.stderr.flush()
                L.lr_tc_debug_stats()
                L.lr_tc_timing_dump()
    engine.match_set_mode(0)


if __name__ == "__main__":
    main()
