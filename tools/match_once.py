"""One cfg-2 mutual-NN match (50k x 50k x 32) repeated a few times: the workload ncu profiles the matching kernels on.
usage: python tools/match_once.py [mode] [reps]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from lidarregistration_b200 import engine  # noqa: E402

mode = int(sys.argv[1]) if len(sys.argv) > 1 else 0
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
n = 50000
dev = torch.device("cuda:0")
g = torch.Generator(device=dev).manual_seed(51 + 2000)
f0 = torch.nn.functional.normalize(torch.randn(n, 32, device=dev, generator=g), dim=1)
f1 = torch.nn.functional.normalize(torch.randn(n, 32, device=dev, generator=g), dim=1)
f1[:n // 2] = torch.nn.functional.normalize(f0[:n // 2] + 0.08 * torch.randn(n // 2, 32, device=dev, generator=g), dim=1)
engine.match_set_mode(mode)
for _ in range(reps):
    i1, _ = engine.match_nn(f0, f1, want_2nd=False)
    mi, mj = engine.match_mutual(f0, f1, i1)
torch.cuda.synchronize()
print("mutual pairs", int(mi.shape[0]))
