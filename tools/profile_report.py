"""Turn ncu captures (read here, no GPU) into the tracked summaries under profiles/.

usage: python tools/profile_report.py <round-tag> <full.ncu-rep> <frames-in-capture> <workload> [<launches.csv>]
writes profiles/<round-tag>_<workload>_ncu_summary.md, updates profiles/r2_traffic.json (DRAM bytes per frame and kernel),
and (with a launch list) profiles/<round-tag>_<workload>_launches.md with each kernel's share of the step."""
import collections
import csv
import io
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KMAP = [("k_sweep", "rx_sweep"), ("k_stats<float, 0", "me_stats"), ("k_stats<unsigned char, 0", "me_stats"),
        ("k_stats<float, 1", "nvf_stats"), ("k_stats<unsigned char, 1", "nvf_stats"),
        ("k_apply<float, float, 0", "me_apply"), ("k_apply<unsigned char, unsigned char, 0", "me_apply"),
        ("k_apply<float, float, 1", "nvf_apply"), ("k_apply<unsigned char, unsigned char, 1", "nvf_apply"),
        ("k_apply_ts<float, 0", "me_apply"), ("k_apply_ts<unsigned char, 0", "me_apply"),
        ("k_apply_ts<float, 1", "nvf_apply"), ("k_apply_ts<unsigned char, 1", "nvf_apply"),
        ("k_detect<float, 0", "me_detect"), ("k_detect<unsigned char, 0", "me_detect"),
        ("k_detect<float, 1", "nvf_detect"), ("k_detect<unsigned char, 1", "nvf_detect")]


def kname(full):
    for pat, n in KMAP:
        if pat in full:
            return n
    return None


def raw(rep):
    # a .csv is the raw page already exported on the GPU box (`ncu -i x.ncu-rep --page raw --csv`): the captures
    # themselves (40-50 MB each) do not fit gpurun's 64 MiB return channel together
    if rep.endswith(".csv"):
        out = open(rep).read()
    else:
        out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    return rows[0], rows[1], rows[2:]


def fnum(x):
    try:
        return float(x.replace(",", ""))
    except ValueError:
        return float("nan")


def to_bytes(v, unit):
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1)


def main():
    tag, rep, frames, wl = sys.argv[1], sys.argv[2], int(sys.argv[3]), sys.argv[4]
    launches = sys.argv[5] if len(sys.argv) > 5 else None
    hdr, units, data = raw(rep)
    ix = {h: i for i, h in enumerate(hdr)}
    cols = [("gpu__time_duration.sum", "time"), ("dram__bytes_read.sum", "DRAM rd"), ("dram__bytes_write.sum", "DRAM wr"),
            ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM %"),
            ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue %"),
            ("sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "FMA pipe %"),
            ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps %"),
            ("launch__registers_per_thread", "regs"), ("launch__grid_size", "grid"), ("smsp__inst_executed.sum", "warp instr")]
    stall = [h for h in hdr if "warps_issue_stalled" in h and h.endswith("per_issue_active.ratio")]
    lines = ["# %s — ncu `--set full --clock-control none` summary, workload `%s` (%d frames per launch)" % (tag, wl, frames), "",
             "Source capture: `%s` (scratch, not tracked). Clocks are not locked under ncu and caches are cold, so absolute "
             "times differ from bench.py's CUDA-event times; DRAM bytes and instruction counts are exact." % os.path.basename(rep), "",
             "| kernel | " + " | ".join(c[1] for c in cols) + " | instr/px | top stalls (per issue) |", "|---|" + "---|" * (len(cols) + 2)]
    traffic = {}
    instr = {}
    seen = set()
    for r in data:
        full = r[ix["Kernel Name"]]
        short = re.sub(r"\(.*", "", full).replace("void ", "")
        if short in seen:
            continue
        seen.add(short)
        vals = []
        for m, _ in cols:
            v, u = r[ix[m]], units[ix[m]]
            vals.append("%s %s" % (v.rstrip("0").rstrip(".") if "." in v else v, u) if u not in ("", "%") else v[:6])
        st = sorted(((fnum(r[ix[h]]), h.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", "")) for h in stall), reverse=True)[:4]
        lines.append("| `%s` | %s | %s |" % (short, " | ".join(vals), ", ".join("%s %.2f" % (n, v) for v, n in st)))
        kn = kname(full)
        if kn:
            b = to_bytes(fnum(r[ix["dram__bytes_read.sum"]]), units[ix["dram__bytes_read.sum"]]) + \
                to_bytes(fnum(r[ix["dram__bytes_write.sum"]]), units[ix["dram__bytes_write.sum"]])
            traffic.setdefault(kn, []).append(b / frames)
            instr.setdefault(kn, []).append(fnum(r[ix["smsp__inst_executed.sum"]]) / frames)
    # instr/px needs the pixel count: filled by the caller's knowledge of the workload
    px = {"image1080p": 1080 * 1920, "single1080p": 1080 * 1920, "video4k": 2160 * 3840, "image4k": 2160 * 3840, "image512": 512 * 512}.get(wl)
    if px:
        out = []
        for ln in lines:
            m = re.match(r"\| `", ln)
            if m:
                parts = ln.split(" | ")
                winst = fnum(parts[len(cols)].split()[0])
                parts[len(cols) + 1] = "%.1f" % (winst * 32 / (px * frames)) + " | " + parts[len(cols) + 1]
                ln = " | ".join(parts)
            out.append(ln)
        lines = out
    os.makedirs(os.path.join(ROOT, "profiles"), exist_ok=True)
    with open(os.path.join(ROOT, "profiles", "%s_%s_ncu_summary.md" % (tag, wl)), "w") as f:
        f.write("\n".join(lines) + "\n")
    tpath = os.path.join(ROOT, "profiles", "r2_traffic.json")
    t = json.load(open(tpath)) if os.path.exists(tpath) else {}
    t[wl] = {"unit": "DRAM bytes (read+write) per frame per launch, from ncu --set full", "source": tag,
             "per_frame": {k: sum(v) / len(v) for k, v in traffic.items()},
             "warp_instr_per_frame": {k: sum(v) / len(v) for k, v in instr.items()}}
    json.dump(t, open(tpath, "w"), indent=1)
    if launches:
        rows = [r for r in csv.reader(open(launches)) if r and not r[0].startswith("==")]
        h = rows[0]
        ki, vi, ui = h.index("Kernel Name"), h.index("Metric Value"), h.index("Metric Unit")
        tot, cnt = collections.Counter(), collections.Counter()
        for r in rows[1:]:
            n = re.sub(r"\(.*", "", r[ki]).replace("void ", "")
            v = fnum(r[vi]) * {"ns": 1e-3, "us": 1, "ms": 1e3}.get(r[ui], 1)
            tot[n] += v
            cnt[n] += 1
        T = sum(tot.values())
        L = ["# %s — launch list of `bench.py` (`ncu --metrics gpu__time_duration.sum --clock-control none`), workload `%s`" % (tag, wl), "",
             "Per-launch times under ncu are cold-cache and serialised: compare SHARES with bench.py's `kernels` list, not absolutes.", "",
             "| kernel | launches | total us | share | avg us |", "|---|---|---|---|---|"]
        for n, v in tot.most_common():
            L.append("| `%s` | %d | %.1f | %.1f %% | %.1f |" % (n, cnt[n], v, 100 * v / T, v / cnt[n]))
        with open(os.path.join(ROOT, "profiles", "%s_%s_launches.md" % (tag, wl)), "w") as f:
            f.write("\n".join(L) + "\n")


if __name__ == "__main__":
    main()
