#!/usr/bin/env python
"""Turns the raw artefacts a gpurun call leaves in gpurun_out/ (bench JSON lines, ncu launch lists, ncu --set full
reports) into the tracked summaries under profiles/.  Usage: python tools/make_profiles.py [round-tag, default r01]"""
import collections
import csv
import json
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G, P = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")
TAG = sys.argv[1] if len(sys.argv) > 1 else "r01"


def agg(path):
    rows = list(csv.reader(open(path)))
    hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    hdr, data = rows[hi], rows[hi + 1:]
    ki, mi, vi, ii = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("ID")
    per = collections.OrderedDict()
    for r in data:
        if len(r) > vi:
            per.setdefault(r[ii], {"k": r[ki]})[r[mi]] = float(r[vi].replace(",", ""))
    a = collections.OrderedDict()
    for v in per.values():
        e = a.setdefault(v["k"], [0, 0.0, 0.0, 0.0])
        e[0] += 1
        e[1] += v.get("gpu__time_duration.sum", 0)
        e[2] += v.get("dram__bytes_read.sum", 0)
        e[3] += v.get("dram__bytes_write.sum", 0)
    return a


def short(k):
    return k.replace("hg::<unnamed>::", "").replace("void ", "").replace("CUB_200802_SM_1000::", "cub::")[:84]


def copy(src, dst):
    if os.path.exists(os.path.join(G, src)):
        shutil.copy(os.path.join(G, src), os.path.join(P, dst))
        return True
    return False


def launches():
    if not copy("launches_r1b.csv", "%s_launches.csv" % TAG):
        return
    a = agg(os.path.join(P, "%s_launches.csv" % TAG))
    tot = sum(v[1] for v in a.values())
    d = json.load(open(os.path.join(P, "%s_bench_ours.json" % TAG)))
    L = ["# Round 1 — ncu launch list of `HG_BENCH_SKIP_TRAIN=1 python bench.py --steps 2 --warmup 3` (B200, 1M Gaussians, 1080p)\n",
         "Command: `ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 600 --csv`",
         "(cold-cache, serialised: compare SHARES).  The capture covers the resident-leg steps and the first steps of the e2e leg",
         "(with torch's elementwise kernels of its L1 loss).\n",
         "| kernel | launches | total us | avg us | share | DRAM read MB/launch | DRAM write MB/launch |", "|---|---:|---:|---:|---:|---:|---:|"]
    for k, v in sorted(a.items(), key=lambda x: -x[1][1])[:22]:
        L.append("| `%s` | %d | %.1f | %.1f | %.3f | %.1f | %.1f |" % (short(k), v[0], v[1] / 1e3, v[1] / 1e3 / v[0], v[1] / tot,
                                                                    v[2] / v[0] / 1e6, v[3] / v[0] / 1e6))
    L.append("\nTotal captured GPU time: %.1f us." % (tot / 1e3))
    st = d["roofline"]["stage_ms"]
    ssum = sum(st.values())
    L.append("\nLive CUDA-event stage times of the same build (`profiles/%s_bench_ours.json`): " % TAG +
             ", ".join("%s %.3f ms (%.1f %%)" % (k, v, 100 * v / ssum) for k, v in st.items()) + ".")
    rast = sum(v[1] for k, v in a.items() if "hg::" in k or "cub" in k.lower())
    sh = lambda sub: sum(v[1] for k, v in a.items() if sub in k) / rast  # noqa: E731
    L.append("Shares inside the rasterizer's own kernels (ncu): blend_bwd %.3f, blend_fwd %.3f, radix sorts %.3f, preprocess_bwd "
             "%.3f, preprocess_fwd %.3f — they agree with the live stage shares above (blend_bwd %.3f, blend_fwd %.3f)." %
             (sh("blend_bwd"), sh("blend_fwd"), sh("RadixSort"), sh("preprocess_bwd"), sh("preprocess_fwd"),
              st["blend_bwd"] / ssum, st["blend_fwd"] / ssum))
    open(os.path.join(P, "%s_launches_summary.md" % TAG), "w").write("\n".join(L) + "\n")
    tr = {}
    tp = os.path.join(P, "ncu_traffic.json")
    if os.path.exists(tp):
        tr = json.load(open(tp))
    for k, v in a.items():
        for name in ("blend_fwd", "blend_bwd", "preprocess_bwd", "preprocess_fwd"):
            if name in k and "_kernel" in k:
                tr[name] = int((v[2] + v[3]) / v[0])
    json.dump(tr, open(tp, "w"), indent=1)


def table(src, dst, title, note, n_iter):
    if not copy(src, dst + ".csv"):
        return
    a = agg(os.path.join(P, dst + ".csv"))
    tot = sum(v[1] for v in a.values())
    L = [title + "\n", note + "\n", "| kernel | launches / iteration | avg us | us / iteration | share |", "|---|---:|---:|---:|---:|"]
    for k, v in sorted(a.items(), key=lambda x: -x[1][1])[:26]:
        L.append("| `%s` | %.1f | %.1f | %.1f | %.3f |" % (short(k), v[0] / n_iter, v[1] / 1e3 / v[0], v[1] / 1e3 / n_iter, v[1] / tot))
    L.append("\nTotal: %.1f us per iteration under ncu (cold, serialised)." % (tot / 1e3 / n_iter))
    open(os.path.join(P, dst + "_summary.md"), "w").write("\n".join(L) + "\n")


def blend_full(rep):
    path = os.path.join(G, rep)
    if not os.path.exists(path):
        return
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
            "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
            "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
            "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
            "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
            "smsp__thread_inst_executed_per_inst_executed.ratio", "sm__warps_active.avg.pct_of_peak_sustained_active",
            "launch__registers_per_thread", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
            "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smsp__inst_executed_op_global_red.sum", "lts__t_sector_hit_rate.pct",
            "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
            "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
            "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
            "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
            "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
            "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
            "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio"]
    kn = hdr.index("Kernel Name")
    kern = rows[2:]
    names = [r[kn].split("(")[0].split("::")[-1] for r in kern]
    L = ["# Round 1 — `ncu --set full --clock-control none --import-source on -k regex:blend_` (B200)\n",
         "Source: gpurun_out/%s (not tracked), one launch of each blend kernel inside `HG_BENCH_SKIP_TRAIN=1 python bench.py --steps 2" % rep,
         "--warmup 3` (1M Gaussians, 1920x1080, R = 5.42 M tile instances, geometry + depth outputs on).\n",
         "| metric | " + " | ".join(names) + " | unit |", "|---|" + "---:|" * len(names) + "---|"]
    vals = {}
    for w in want:
        if w in hdr:
            i = hdr.index(w)
            L.append("| %s | %s | %s |" % (w, " | ".join(r[i] for r in kern), units[i]))
            vals[w] = [r[i] for r in kern]
    open(os.path.join(P, "%s_ncu_blend_full.md" % TAG), "w").write("\n".join(L) + "\n\n" + open(os.path.join(P, "%s_ncu_blend_reading.md" % TAG)).read()
                                                                  if os.path.exists(os.path.join(P, "%s_ncu_blend_reading.md" % TAG)) else "\n".join(L) + "\n")
    tp = os.path.join(P, "ncu_traffic.json")
    tr = json.load(open(tp)) if os.path.exists(tp) else {}
    key = lambda n: "blend_fwd" if "fwd" in n else "blend_bwd"  # noqa: E731
    tr["_issue_active_pct"] = {key(n): float(v) for n, v in zip(names, vals["smsp__issue_active.avg.pct_of_peak_sustained_active"])}
    tr["_warp_instructions"] = {key(n): float(v) for n, v in zip(names, vals["smsp__inst_executed.sum"])}
    json.dump(tr, open(tp, "w"), indent=1)


def main():
    os.makedirs(P, exist_ok=True)
    copy("bench_r1_final.json", "%s_bench_ours.json" % TAG)
    copy("bench_r1_final_ref.json", "%s_bench_reference.json" % TAG)
    copy("bench_n2b.json", "%s_bench_ours_n2.json" % TAG)
    launches()
    table("loss_launches2.csv", "%s_loss_launches" % TAG,
          "# Round 1 — ncu launch list of the loss path: `python tools/loss_bench.py --no-cpu --iters 3 --warmup 2` (B200)",
          "One iteration = L1 + SSIM + frequency_regularization_pyramid_scale forward + backward on a 3x1080x1920 pair (BASELINE "
          "configs[0]); the capture holds 11 iterations.", 11)
    table("train_launches2.csv", "%s_train_launches" % TAG,
          "# Round 1 — ncu launch list of the training step: `python tools/train_probe.py --recipe c2 --steps 2` (B200)",
          "Config 3 (config-2 scene at 2M Gaussians, 1080p, one view per step: render + all losses + backward + Adam); the capture "
          "starts after 400 launches and holds about 2.5 steps.", 2.5)
    blend_full("prof_blend_r1c.ncu-rep")


if __name__ == "__main__":
    main()
