"""Turn gpurun_out ncu artefacts into small text summaries under profiles/ (tracked).
  python tools/summarize_ncu.py launches gpurun_out/launches.csv profiles/r01_launches.txt
  python tools/summarize_ncu.py raw gpurun_out/prof_x.ncu-rep profiles/r01_prof_x.txt
"""
import collections
import csv
import re
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tensor_op_utcmma.sum", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__waves_per_multiprocessor",
        "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active", "lts__t_sector_hit_rate.pct",
        "launch__shared_mem_per_block_dynamic"]


def clean(name):
    return re.sub(r"\(.*", "", name).replace("void ", "").replace("dmc::<unnamed>::", "")


def launches(src, dst):
    rows = list(csv.reader(open(src)))
    h = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    hdr, data = rows[h], rows[h + 1:]
    iname, ival, iunit = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    ig = hdr.index("Grid Size")
    seq = []
    for r in data:
        if len(r) <= ival:
            continue
        v = float(r[ival].replace(",", ""))
        v = v / 1000.0 if r[iunit] == "ns" else v
        seq.append((clean(r[iname]), v, r[ig]))
    idx = [i for i, (n, _, _) in enumerate(seq) if n.startswith("ema_kernel") or n.startswith("ema2_kernel")]
    a, b = idx[-3] + 1, idx[-2] + 1
    step = seq[a:b]
    tot = sum(v for _, v, _ in step)
    out = [f"# ncu --metrics gpu__time_duration.sum --clock-control none: one whole step ({len(step)} launches)",
           f"# per-launch times are cold-cache and serialised; compare SHARES.  sum = {tot:.1f} us", ""]
    agg, cnt = collections.Counter(), collections.Counter()
    for n, v, _ in step:
        agg[n] += v
        cnt[n] += 1
    out.append("## by kernel")
    for n, v in agg.most_common():
        out.append(f"{n[:72]:72s} x{cnt[n]:<2d} {v:8.1f} us {100 * v / tot:5.1f}%")
    out += ["", "## launch order"]
    for n, v, g in step:
        out.append(f"{n[:72]:72s} {v:8.1f} us  grid={g}")
    open(dst, "w").write("\n".join(out) + "\n")
    print("\n".join(out[:30]))


def raw(src, dst):
    txt = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    hdr, units = rows[0], rows[1]
    out = [f"# ncu --set full --clock-control none ({src})", ""]
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        out.append("## " + clean(d.get("Kernel Name", "?")))
        for k in KEYS:
            if k in d:
                out.append(f"  {k:64s} {d[k]:>16s} {units[hdr.index(k)]}")
        out.append("")
    open(dst, "w").write("\n".join(out) + "\n")
    print("\n".join(out[:60]))


if __name__ == "__main__":
    {"launches": launches, "raw": raw}[sys.argv[1]](sys.argv[2], sys.argv[3])
