"""Cross-compile A/B variants of liblidarreg.so HERE (no GPU needed); the files travel to the GPU box with the snapshot
(*.so is git-ignored, not gpurun-ignored) and are selected there with LIDARREG_SO=<path>.
usage: python tools/build_variants.py name1="-DX=1 -DY=0" name2="..."   -> lidarregistration_b200/csrc/variants/lib_<name>.so"""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
VDIR = os.path.join(ROOT, "lidarregistration_b200", "csrc", "variants")
os.makedirs(VDIR, exist_ok=True)
procs = []
for spec in sys.argv[1:]:
    name, _, flags = spec.partition("=")
    so = os.path.join(VDIR, "lib_%s.so" % name)
    env = dict(os.environ, LIDARREG_SO=so, LIDARREG_NVCC_FLAGS=flags)
    procs.append((name, so, subprocess.Popen([sys.executable, "-m", "lidarregistration_b200.build", "--force"], cwd=ROOT, env=env,
                                             stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
for name, so, p in procs:
    out, _ = p.communicate()
    print(name, "OK" if p.returncode == 0 else "FAILED\n" + out[-800:], so)
