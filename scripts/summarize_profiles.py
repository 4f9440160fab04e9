"""Turn the raw captures of scripts/capture_profiles.sh (gpurun_out/<tag>_*) into the tracked summaries under
profiles/ (needs only the ncu CLI, no GPU):

  python scripts/summarize_profiles.py r02

  <tag>_launches.csv   -> profiles/<tag>_ncu_launch_list.csv (verbatim) + profiles/<tag>_launch_summary.txt
                          (per kernel: launches and time of ONE step = the launches between the last two
                          prep_scalars_kernel launches, the first kernel of a forward)
  <tag>_tiles.ncu-rep  -> profiles/<tag>_ncu_full_tile_kernels.csv (selected metrics of every captured launch)
  <tag>_bench_*.json   -> profiles/ (verbatim)
"""
import collections
import csv
import io
import os
import re
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "gpurun_out")
PROF = os.path.join(ROOT, "profiles")

METRICS = [
    ("gpu__time_duration.sum", 1e-6, "ms"),
    ("dram__bytes_read.sum", 1e-9, "Gbyte"),
    ("dram__bytes_write.sum", 1e-9, "Gbyte"),
    ("lts__t_bytes.sum", 1e-9, "Gbyte"),
    ("sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active", 1, "%"),
    ("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", 1, "%"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", 1, "%"),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", 1, "%"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", 1, "%"),
    ("launch__registers_per_thread", 1, "register/thread"),
    ("launch__grid_size", 1, ""),
    ("launch__cluster_max_active", 1, "cluster"),
    ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", 1, "inst"),
    ("smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", 1, "inst"),
    ("smsp__average_warps_issue_stalled_membar_per_issue_active.ratio", 1, "inst"),
    ("smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", 1, "inst"),
    ("smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", 1, "inst"),
]
UNIT_SCALE = {"ns": 1.0, "us": 1e3, "ms": 1e6, "s": 1e9, "byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}


def short(name):
    return re.sub(r"\(.*", "", name).replace("dsoft::", "").strip()


def launch_summary(tag):
    src = os.path.join(OUT, f"{tag}_launches.csv")
    if not os.path.exists(src):
        return
    shutil.copy(src, os.path.join(PROF, f"{tag}_ncu_launch_list.csv"))
    with open(src) as f:
        rows = list(csv.DictReader(l for l in f if not l.startswith("==")))
    starts = [i for i, r in enumerate(rows) if "prep_scalars_kernel" in r["Kernel Name"]]
    if len(starts) < 2:
        return
    # a step = forward ... backward ... up to the next forward's first kernel; take the last complete one
    step = rows[starts[-2]:starts[-1]]
    agg = collections.OrderedDict()
    for r in step:
        k = short(r["Kernel Name"])[:84]
        ns = float(r["Metric Value"]) * UNIT_SCALE.get(r["Metric Unit"], 1.0)
        a = agg.setdefault(k, [0, 0.0])
        a[0] += 1
        a[1] += ns
    tot = sum(v[1] for v in agg.values())
    ours = sum(v[1] for k, v in agg.items() if "at::" not in k and "nvjet" not in k and "cublas" not in k)
    with open(os.path.join(PROF, f"{tag}_launch_summary.txt"), "w") as f:
        f.write("One fwd+bwd step at B=32768 on one B200 (bench.py --steps 2 --warmup 3 under\n"
                "ncu --metrics gpu__time_duration.sum --clock-control none; per-launch times are cold-cache and\n"
                f"serialised: the SHARE of the step is the meaningful number).  Full list: {tag}_ncu_launch_list.csv\n\n")
        f.write("        us   share   n  kernel\n")
        for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f"{v[1] / 1e3:10.1f}  {100 * v[1] / tot:5.1f}%  {v[0]:2d}  {k}\n")
        f.write(f"\nstep total {tot / 1e6:.3f} ms in {len(step)} launches; kernels of libdsoft.so: "
                f"{100 * ours / tot:.1f} % of the step (the rest: cuBLAS GEMMs of the MLP head and small torch "
                "element-wise kernels)\n")


def full_summary(tag, rep=None, out_name=None):
    rep = rep or os.path.join(OUT, f"{tag}_tiles.ncu-rep")
    if not os.path.exists(rep):
        return
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    col = {h: i for i, h in enumerate(hdr)}
    kn = col["Kernel Name"]
    out = os.path.join(PROF, out_name or f"{tag}_ncu_full_tile_kernels.csv")
    with open(out, "w", newline="") as f:
        f.write('"# ncu --set full --clock-control none, bench.py --steps 2 --warmup 3 (B=32768, one B200), one '
                'launch of each kernel of libdsoft.so in a step; units: ' +
                ", ".join(f"{m}={u}" for m, _, u in METRICS) + '"\n')
        w = csv.writer(f)
        w.writerow(["id", "kernel"] + [m for m, _, _ in METRICS])
        for r in data:
            vals = []
            for m, _, want in METRICS:
                if m not in col:
                    vals.append("")
                    continue
                raw = r[col[m]].replace(",", "")
                try:
                    x = float(raw)
                except ValueError:
                    vals.append(raw)
                    continue
                have = units[col[m]]
                if want in ("ms", "Gbyte") and have in UNIT_SCALE:
                    x = x * UNIT_SCALE[have] / UNIT_SCALE[want if want != "ms" else "ms"]
                vals.append(f"{x:.6g}")
            w.writerow([r[col["ID"]], short(r[kn])] + vals)
    return out


def traffic_table(tag, summary=None):
    """profiles/ncu_traffic.csv (read by bench.py for `roofline.traffic`): DRAM bytes read + written per launch of
    the default workload's tile kernels, keyed by bench.py's kernel names.  One step at world 1 launches, in order:
    soft forward <13> (32x32b form: <9>), CLIP forward <10>, soft G <4>, student GEMM, text GEMM, CLIP G <3>, image GEMM, text^T GEMM."""
    summary = summary or os.path.join(PROF, f"{tag}_ncu_full_tile_kernels.csv")
    if not os.path.exists(summary):
        return
    with open(summary) as f:
        rows = list(csv.DictReader(l for l in f if not l.startswith('"#')))
    names = {"dsoft_fwd_kernel<9, 2>": ["fwd_soft"], "dsoft_fwd_kernel<13, 2>": ["fwd_soft"],
             "dsoft_fwd_kernel<10, 2>": ["fwd_clip_i2t"],
             "dsoft_fwd_kernel<4, 2>": ["bwd_build_g_soft"], "dsoft_fwd_kernel<3, 2>": ["bwd_build_g_clip"],
             "dsoft_gy_kernel<0>": ["bwd_student", "bwd_text", "bwd_clip_image"], "dsoft_gy_kernel<1>": ["bwd_clip_text"]}
    sha = subprocess.run(["git", "-C", ROOT, "rev-parse", "--short", "HEAD"], capture_output=True, text=True).stdout.strip()
    seen = collections.Counter()
    with open(os.path.join(PROF, "ncu_traffic.csv"), "w") as f:
        f.write("# DRAM bytes (dram__bytes_read.sum + dram__bytes_write.sum) per launch, ncu --set full --clock-control none,\n"
                "# default bench workload (B=32768, one B200); written by scripts/summarize_profiles.py\n")
        f.write("kernel,dram_bytes,source,git_sha\n")
        for r in rows:
            k = r["kernel"].replace("void ", "")
            if k not in names or seen[k] >= len(names[k]):
                continue
            name = names[k][seen[k]]
            seen[k] += 1
            tot = (float(r["dram__bytes_read.sum"]) + float(r["dram__bytes_write.sum"])) * 1e9
            f.write(f"{name},{tot:.6g},{os.path.basename(summary)},{sha}\n")


def main():
    tag = sys.argv[1]
    os.makedirs(PROF, exist_ok=True)
    launch_summary(tag)
    full_summary(tag)
    traffic_table(tag)
    for fn in sorted(os.listdir(OUT)):
        if fn.startswith(f"{tag}_bench") and fn.endswith(".json"):
            shutil.copy(os.path.join(OUT, fn), os.path.join(PROF, fn))
    print("wrote profiles/%s_*" % tag)


if __name__ == "__main__":
    main()
