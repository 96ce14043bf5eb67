"""Summarise an .ncu-rep (read here, no GPU): per-kernel time, DRAM bytes, issue utilisation, top stalls, opcode mix.
usage: python tools/ncu_summary.py gpurun_out/prof.ncu-rep [--ops KERNEL_REGEX]"""
import collections
import csv
import io
import subprocess
import sys


def raw(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    return rows[0], rows[1], rows[2:]


def main():
    rep = sys.argv[1]
    hdr, units, data = raw(rep)
    idx = {h: i for i, h in enumerate(hdr)}
    want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
            "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
            "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
            "launch__registers_per_thread", "launch__grid_size", "smsp__inst_executed.sum", "sm__cycles_elapsed.avg",
            "sm__cycles_active.avg", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
            "sm__inst_executed_pipe_fma.sum", "sm__inst_executed_pipe_alu.sum", "sm__inst_executed_pipe_fp16.sum",
            "sm__inst_executed_pipe_lsu.sum", "sm__inst_executed_pipe_xu.sum", "sm__inst_executed_pipe_fmaheavy.sum",
            "sm__inst_executed_pipe_fmalite.sum", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
            "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
            "sm__pipe_fp16_cycles_active.avg.pct_of_peak_sustained_active", "lts__t_sector_hit_rate.pct"]
    stall = [h for h in hdr if "warps_issue_stalled" in h and h.endswith("per_issue_active.ratio")]
    for r in data:
        print("----", r[idx["Kernel Name"]][:70])
        for w in want:
            if w in idx:
                print("   %-70s %s %s" % (w, r[idx[w]], units[idx[w]]))
        vals = []
        for h in stall:
            try:
                vals.append((float(r[idx[h]]), h.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", "")))
            except ValueError:
                pass
        print("   stalls/issue:", ", ".join("%s %.2f" % (h, v) for v, h in sorted(vals, reverse=True)[:7]))
    if "--ops" in sys.argv:
        pat = sys.argv[sys.argv.index("--ops") + 1]
        out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + pat,
                              "--launch-count", "1"], capture_output=True, text=True).stdout
        rows = list(csv.reader(io.StringIO(out)))
        hi = next(i for i, r in enumerate(rows) if "Source" in r and "Instructions Executed" in r)
        h = rows[hi]
        si, ei = h.index("Source"), h.index("Instructions Executed")
        cnt, tot = collections.Counter(), 0
        for r in rows[hi + 1:]:
            try:
                n = int(r[ei])
            except (ValueError, IndexError):
                continue
            t = r[si].split()
            op = (t[1] if t[0].startswith("@") else t[0]).split(".")[0]
            cnt[op] += n
            tot += n
        print("opcode mix of", pat, "total warp-instr", tot)
        for op, n in cnt.most_common(22):
            print("   %10d %5.1f%% %s" % (n, 100.0 * n / tot, op))


if __name__ == "__main__":
    main()
