#!/usr/bin/env python
"""Turn ncu artefacts brought back in gpurun_out/ into the text summaries committed under profiles/.

  python tools/summarize_profiles.py launches gpurun_out/launches_r1.csv profiles/r1_launches.txt
  python tools/summarize_profiles.py kernel   gpurun_out/prof_score_r1.ncu-rep profiles/r1_ncu_k_score.txt
"""
import collections
import csv
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "sm__cycles_elapsed.avg.per_second", "launch__grid_size", "launch__block_size",
        "launch__registers_per_thread", "launch__shared_mem_per_block_static", "launch__shared_mem_per_block_dynamic",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__inst_executed.avg.per_cycle_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_subpipe_hmma_cycles_active", "sm__inst_executed_pipe_tensor",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_bytes.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "smsp__average_warps_issue_stalled"]


def launches(src, dst):
    rows = [r for r in csv.reader(open(src)) if len(r) > 10]
    h = rows[0]
    ki, vi, gi, bi = h.index("Kernel Name"), h.index("Metric Value"), h.index("Grid Size"), h.index("Block Size")
    agg = collections.OrderedDict()
    for r in rows[1:]:
        try:
            v = float(r[vi].replace(",", ""))
        except ValueError:
            continue
        name = r[ki].split("(")[0].replace("<unnamed>::", "").replace("void ", "")
        a = agg.setdefault(name, [0, 0.0, r[gi], r[bi]])
        a[0] += 1
        a[1] += v
    tot = sum(a[1] for a in agg.values())
    with open(dst, "w") as f:
        f.write("# ncu --metrics gpu__time_duration.sum --clock-control none, per-launch times are cold-cache and\n"
                "# serialised: compare SHARES, not absolutes.  source: %s\n" % src)
        f.write("%-44s %7s %12s %8s  %s\n" % ("kernel", "launches", "total us", "share", "grid x block (last)"))
        for name, (c, t, g, b) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write("%-44s %7d %12.1f %7.1f%%  %s x %s\n" % (name[:44], c, t / 1e3, 100 * t / tot, g, b))
    print(open(dst).read())


def kernel(src, dst):
    """one summary per profiled launch (distinct kernel names) of the report"""
    raw = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    h, units = rows[0], rows[1]
    seen = set()
    with open(dst, "w") as f:
        f.write("# ncu --set full --clock-control none --import-source on.  source: %s\n" % src)
        for vals in rows[2:]:
            name = vals[h.index("Kernel Name")]
            if name in seen:
                continue
            seen.add(name)
            f.write("\nkernel: %s\n" % name)
            for i, n in enumerate(h):
                if any(n.startswith(k) for k in KEYS) and ".min" not in n and ".max" not in n and ".sum.p" not in n \
                        and "peak_sustained_elapsed" not in n.replace("sm__throughput", "").replace("gpu__dram", ""):
                    f.write("%-86s %s %s\n" % (n, vals[i], units[i]))
            short = name.split("(")[0].split("::")[-1].replace("void ", "").split("<")[0]
            # top stall sites from the source page of this kernel
            srcp = subprocess.run(["ncu", "-i", src, "--page", "source", "--csv", "--kernel-name", "regex:" + short],
                                  capture_output=True, text=True).stdout
            srows = list(csv.reader(srcp.splitlines()))
            sh = None
            for k, r in enumerate(srows[:6]):
                if "Source" in r and "# Samples" in r:
                    sh, first = r, k + 1
                    break
            if sh is None:
                continue
            isrc, ismp, iex = sh.index("Source"), sh.index("# Samples"), sh.index("Instructions Executed")
            names = [n for n in sh if n.startswith("stall_") and "Not Issued" not in n]
            data = []
            for r in srows[first:]:
                if r and r[0] == sh[0]:
                    break  # the next launch's listing
                try:
                    data.append((int(r[ismp]), int(r[iex]), r[isrc], r))
                except (ValueError, IndexError):
                    pass
            tot = sum(d[0] for d in data) or 1
            agg = collections.Counter()
            for d in data:
                for n in names:
                    v = d[3][sh.index(n)]
                    if v not in ("0", ""):
                        agg[n] += int(v)
            f.write("warp-stall samples by reason (all warps): " +
                    ", ".join("%s %.1f%%" % (k[6:], 100 * v / tot) for k, v in agg.most_common(8)) + "\n")
            f.write("top sampled SASS instructions:\n")
            for d in sorted(data, key=lambda x: -x[0])[:12]:
                f.write("  %5.1f%%  executed %10d  %s\n" % (100 * d[0] / tot, d[1], d[2][:90]))
    print(open(dst).read())


if __name__ == "__main__":
    {"launches": launches, "kernel": kernel}[sys.argv[1]](sys.argv[2], sys.argv[3])
