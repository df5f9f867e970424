"""find_nn / mutual match wall time on the cfg-2 workload (50k x 50k x 32) with programmatic dependent launch off / on."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from lidarregistration_b200 import _lib, engine  # noqa: E402

n = 50000
dev = torch.device("cuda:0")
g = torch.Generator(device=dev).manual_seed(51 + 2000)
f0 = torch.nn.functional.normalize(torch.randn(n, 32, device=dev, generator=g), dim=1)
f1 = torch.nn.functional.normalize(torch.randn(n, 32, device=dev, generator=g), dim=1)
f1[: n // 2] = torch.nn.functional.normalize(f0[: n // 2] + 0.08 * torch.randn(n // 2, 32, device=dev, generator=g), dim=1)
L = _lib.lib()
out = {}
ref = None
for pdl in (0, 1, 0, 1):
    _lib.check(L.lr_debug_pdl(pdl), "pdl")
    for _ in range(5):
        i1, i2 = engine.match_nn(f0, f1, want_2nd=True)
        mi, mj = engine.match_mutual(f0, f1, i1)
    e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    reps = 30
    t_nn = t_mut = 0.0
    for _ in range(reps):
        e[0].record()
        i1, i2 = engine.match_nn(f0, f1, want_2nd=False)
        e[1].record()
        mi, mj = engine.match_mutual(f0, f1, i1)
        e[2].record()
        torch.cuda.synchronize()
        t_nn += e[0].elapsed_time(e[1]) / reps
        t_mut += e[1].elapsed_time(e[2]) / reps
    sig = (int(i1.sum().item()), int(mi.shape[0]), int(mj.sum().item()))
    ref = ref or sig
    assert sig == ref
    out.setdefault("pdl%d" % pdl, []).append(dict(find_nn_ms=round(t_nn, 4), nn_to_mutual_ms=round(t_mut, 4)))
_lib.check(L.lr_debug_pdl(1), "pdl")
print(json.dumps(out))
