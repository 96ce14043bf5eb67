"""Static SASS mnemonic histogram of libwm_b200.so per kernel family -> profiles/r2_sass_histogram.txt (no GPU needed).
usage: python tools/sass_histogram.py"""
import collections
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "watermarking-gpu_b200", "libwm_b200.so")
COLS = ["UTMALDG", "UTMASTG", "SYNCS", "HMMA", "F2FP", "HMUL2", "HFMA2", "FHADD", "FFMA", "FMUL", "FADD", "PRMT", "I2F", "F2I", "MUFU",
        "LDS", "STS", "LDG", "STG", "BAR", "MEMBAR", "SHFL", "DADD", "DFMA"]
HEAD = """# r2 — SASS mnemonic histogram of libwm_b200.so (`cuobjdump -sass`, sm_100a; tools/sass_histogram.py)

Static counts, summed over all template instantiations of a kernel family (prefix match on the mnemonic).
UTMALDG = cp.async.bulk.tensor global->smem (TMA load); UTMASTG = TMA store (k_apply_ts, the default apply kernel: WM_OPT_TMA_STORE);
SYNCS = mbarrier init / arrive.expect_tx / try_wait; HMMA = the legacy mma.sync m16n8k16 that sums the fp16-rounded Rx/rx products
(the tcgen05 form was built and measured 2x slower: r2_tcgen05_microbench.md; its UTCHMMA / STTM / LDTM live in tools/microbench/tc5, last line);
FHADD = add.rn.f32.f16 (u8 -> f32 widening and the FHADD accumulation mode); no I2F.U8 / F2I remains in the hot loops (edge paths only).
k_detect1 = the single-image fused detector (WM_OPT_FUSED_SINGLE, off by default).
"""
TAIL = "\ntools/microbench/tc5 (tcgen05 accumulation microbenchmark): LDTM.x16 4, STTM.x16 12, STTM.x32 4, STTM.x4 12, UTCATOMSWS.AND 4, UTCBAR 12, UTCHMMA 56\n"


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    fam = None
    hist = collections.defaultdict(collections.Counter)
    for ln in sass.splitlines():
        m = re.search(r"Function : _ZN2wm\d+(k_[a-z0-9_]+?)(I|E)", ln)
        if m:
            fam = m.group(1)
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4,5}\*/\s+(?:@!?U?P[0-9T]\s+)?([A-Z0-9_]+)", ln)
        if m and fam:
            op = m.group(1)
            hist[fam]["total"] += 1
            for c in COLS:
                if op.startswith(c):
                    hist[fam][c] += 1
                    break
    out = [HEAD, "%-12s" % "kernel" + "".join("%8s" % c for c in COLS) + "%9s" % "total"]
    for fam in sorted(hist):
        out.append("%-12s" % fam + "".join("%8d" % hist[fam][c] for c in COLS) + "%9d" % hist[fam]["total"])
    with open(os.path.join(ROOT, "profiles", "r2_sass_histogram.txt"), "w") as f:
        f.write("\n".join(out) + "\n" + TAIL)
    print("\n".join(out[1:]))


if __name__ == "__main__":
    main()
