#!/usr/bin/env python
"""Turns the raw artefacts a gpurun call leaves in gpurun_out/ (bench JSON lines, ncu launch lists, ncu --set full
reports) into the tracked summaries under profiles/.  Usage: python tools/make_profiles.py [round-tag, default r01]"""
import collections
import csv
import io
import json
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G, P = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")
TAG = sys.argv[1] if len(sys.argv) > 1 else "r02"
ROUND = "Round %d" % int(TAG[1:])


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
    if not copy("%s_launches.csv" % TAG, "%s_launches.csv" % TAG):
        return
    a = agg(os.path.join(P, "%s_launches.csv" % TAG))
    tot = sum(v[1] for v in a.values())
    d = json.load(open(os.path.join(P, "%s_bench_ours.json" % TAG)))
    L = ["# %s — ncu launch list of `HG_BENCH_SKIP_TRAIN=1 python bench.py --steps 2 --warmup 3` (B200, 1M Gaussians, 1080p)\n" % ROUND,
         "Command: `ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 400 --csv`",
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
    L.append("Shares inside the rasterizer's own kernels (ncu): blend_bwd %.3f, blend_fwd %.3f, tile_sort %.3f, scatter %.3f, "
             "tile_scan %.3f, preprocess_bwd %.3f, preprocess_fwd %.3f — they agree with the live stage shares above (blend_bwd "
             "%.3f, blend_fwd %.3f).  Library kernels among the rasterizer's launches: %s." %
             (sh("blend_bwd"), sh("blend_fwd"), sh("tile_sort"), sh("scatter_instances"), sh("tile_scan"), sh("preprocess_bwd"),
              sh("preprocess_fwd"), st["blend_bwd"] / ssum, st["blend_fwd"] / ssum,
              ", ".join(sorted(set(short(k)[:40] for k in a if "cub" in k.lower()))) or "none (no cub:: symbol)"))
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


def train_table(src, dst, title, note, views):
    """The LAST optimiser step of the capture (between the two last groups of Adam launches), per view."""
    if not copy(src, dst + ".csv"):
        return
    rows = list(csv.reader(open(os.path.join(P, dst + ".csv"))))
    hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    hdr = rows[hi]
    kn, mv, mn = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Name")
    L_ = [(r[kn], float(r[mv].replace(",", ""))) for r in rows[hi + 1:] if len(r) > mv and r[mn] == "gpu__time_duration.sum"]
    idx = [i for i, (n, _) in enumerate(L_) if "adam_masked4" in n]
    groups = []
    for i in idx:
        if groups and i - groups[-1][-1] < 12:
            groups[-1].append(i)
        else:
            groups.append([i])
    if len(groups) < 2:
        return
    a, b = groups[-2][-1] + 1, groups[-1][-1] + 1
    while a < len(L_) and "adam" in L_[a][0]:
        a += 1
    while b < len(L_) and "adam" in L_[b][0]:
        b += 1
    seg = L_[a:b]
    acc = {}
    for n, t in seg:
        e = acc.setdefault(n, [0, 0.0])
        e[0] += 1
        e[1] += t
    tot = sum(v[1] for v in acc.values())
    out = [title + "\n", note + "\n", "| kernel | launches / step | us / step | us / view | share |", "|---|---:|---:|---:|---:|"]
    for k, v in sorted(acc.items(), key=lambda x: -x[1][1])[:40]:
        out.append("| `%s` | %d | %.1f | %.1f | %.3f |" % (short(k), v[0], v[1] / 1e3, v[1] / 1e3 / views, v[1] / tot))
    out.append("\nTotal: %.1f us per step = %.1f us per view under ncu (cold, serialised), %d launches per step." %
               (tot / 1e3, tot / 1e3 / views, len(seg)))
    open(os.path.join(P, dst + "_summary.md"), "w").write("\n".join(out) + "\n")


def full_capture(rep, dst, what, cmd, traffic_keys=False):
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
            "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
            "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
            "smsp__inst_executed_op_global_red.sum", "lts__t_sector_hit_rate.pct",
            "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
            "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
            "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
            "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
            "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
            "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
            "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio"]
    kn = hdr.index("Kernel Name")
    kern = rows[2:]
    names = [r[kn].split("(")[0].split("::")[-1].replace("void ", "")[:28] for r in kern]
    L = ["# %s — `ncu --set full --clock-control none --import-source on` (B200): %s\n" % (ROUND, what),
         "Source: gpurun_out/%s (not tracked), one launch of each kernel inside `%s`.\n" % (rep, cmd),
         "| metric | " + " | ".join(names) + " | unit |", "|---|" + "---:|" * len(names) + "---|"]
    vals = {}
    for w in want:
        if w in hdr:
            i = hdr.index(w)
            L.append("| %s | %s | %s |" % (w, " | ".join(r[i] for r in kern), units[i]))
            vals[w] = [r[i] for r in kern]
    extra = os.path.join(P, dst.replace(".md", "_reading.md"))
    body = "\n".join(L) + "\n"
    if os.path.exists(extra):
        body += "\n" + open(extra).read()
    open(os.path.join(P, dst), "w").write(body)
    if not traffic_keys:
        return
    tp = os.path.join(P, "ncu_traffic.json")
    tr = json.load(open(tp)) if os.path.exists(tp) else {}
    ia, wi = tr.get("_issue_active_pct", {}), tr.get("_warp_instructions", {})
    for n, a_, w_ in zip(names, vals["smsp__issue_active.avg.pct_of_peak_sustained_active"], vals["smsp__inst_executed.sum"]):
        for key in ("blend_fwd", "blend_bwd"):
            if key in n:
                ia[key], wi[key] = float(a_), float(w_.replace(",", ""))
    tr["_issue_active_pct"], tr["_warp_instructions"] = ia, wi
    # packed FP32 (FADD2 / FMUL2 / FFMA2) instructions executed, from the report's SASS page: each retires two FP32
    # operations in one issue slot, so executed + packed = the scalar-equivalent instruction count of the kernel
    pk = tr.get("_packed_fp32_instructions", {})
    for key, rx in (("blend_fwd", "blend_fwd2"), ("blend_bwd", "blend_bwd3")):
        out = subprocess.run(["ncu", "-i", os.path.join(G, rep), "--page", "source", "--csv", "--kernel-name", "regex:" + rx],
                             capture_output=True, text=True).stdout
        rows = list(csv.reader(io.StringIO(out)))
        hi = [i for i, r in enumerate(rows) if r and r[0] == "Address"]
        if not hi:
            continue
        hdr = rows[hi[0]]
        end = hi[1] - 1 if len(hi) > 1 else len(rows)
        si, ie = hdr.index("Source"), hdr.index("Instructions Executed")
        pk[key] = float(sum(int(r[ie]) for r in rows[hi[0] + 1:end]
                            if len(r) == len(hdr) and any(m in r[si] for m in ("FFMA2", "FMUL2", "FADD2"))))
    tr["_packed_fp32_instructions"] = pk
    json.dump(tr, open(tp, "w"), indent=1)


def main():
    os.makedirs(P, exist_ok=True)
    copy("%s_bench.json" % TAG, "%s_bench_ours.json" % TAG)
    copy("%s_bench_ref.json" % TAG, "%s_bench_reference.json" % TAG)
    copy("%s_hier_probe.json" % TAG, "%s_hier_probe.json" % TAG)
    for n in (2, 4, 8):
        copy("%s_n%d.json" % (TAG, n), "%s_bench_ours_n%d.json" % (TAG, n))
    launches()
    table("%s_loss_launches.csv" % TAG, "%s_loss_launches" % TAG,
          "# %s — ncu launch list of the loss path: `python tools/loss_bench.py --no-cpu --iters 3 --warmup 2` (B200)" % ROUND,
          "One iteration = L1 + SSIM + frequency_regularization_pyramid_scale forward + backward on a 3x1080x1920 pair (BASELINE "
          "configs[0]); the capture holds 11 iterations (full path 5 + per-function passes).", 11)
    train_table("%s_train_launches.csv" % TAG, "%s_train_launches" % TAG,
                "# %s — ncu launch list of the training leg: `python tools/train_leg_probe.py --steps 2` (B200)" % ROUND,
                "bench.py's `train` leg alone (configs[4] recipe: 2M-Gaussian UAV slab in Morton order, 8 views per step, "
                "1920x1080, render + L1 + SSIM + frequency / scale regulariser + normal term + backward per view, then the "
                "sparse Adam); the table is the LAST optimiser step of the capture.", 8)
    full_capture("%s_prof_raster.ncu-rep" % TAG, "%s_ncu_raster_full.md" % TAG, "the rasterizer's kernels",
                 "HG_BENCH_SKIP_TRAIN=1 python bench.py --steps 2 --warmup 3` (1M Gaussians, 1920x1080, R = 5.42 M tile instances, "
                 "geometry + depth outputs on", traffic_keys=True)
    full_capture("%s_prof_loss.ncu-rep" % TAG, "%s_ncu_loss_full.md" % TAG, "the loss path's heaviest kernels",
                 "python tools/loss_bench.py --no-cpu --iters 2 --warmup 1` (3x1080x1920")
    full_capture("%s_prof_uav.ncu-rep" % TAG, "%s_ncu_uav_view_full.md" % TAG,
                 "the rasterizer's kernels on one UAV training view (configs[4] recipe: 2M-Gaussian slab in Morton order, "
                 "~20 % of it rendered, R ~ 0.76 M tile instances)",
                 "python tools/train_leg_probe.py --steps 1` (`-k regex:blend_|preprocess_|tile_sort|scatter -s 60 -c 7")


if __name__ == "__main__":
    main()
