"""The reference's own timing protocol (main.cpp:167-223): one image, 2 warm-ups, `loops` synchronous calls per op,
mean seconds -> FPS.  usage: python tools/latency.py [--rows 1080 --cols 1920 --loops 1000]"""
import argparse
import importlib
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import util  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rows", type=int, default=1080)
    ap.add_argument("--cols", type=int, default=1920)
    ap.add_argument("--loops", type=int, default=1000)
    ap.add_argument("--fused", action="store_true", help="WM_OPT_FUSED_SINGLE = 1: wm_detect as one cooperative kernel (A/B; default off)")
    a = ap.parse_args()
    pkg = importlib.import_module("watermarking-gpu_b200")
    img = util.natural_image(a.rows, a.cols, seed=1)
    W = util.normal_w(a.rows, a.cols)
    wm = pkg.Watermark(a.rows, a.cols, W, 3, 40.0)
    if a.fused:
        wm.set_option(pkg.OPT_FUSED_SINGLE, 1)
    d = pkg.DeviceArray.from_numpy(wm, img, pkg.COL_MAJOR)
    out = pkg.DeviceArray(wm, a.rows, a.cols, pkg.COL_MAJOR, pkg.F32)
    res = {}
    for name, mask in (("NVF", pkg.NVF), ("ME", pkg.ME)):
        for _ in range(2):
            wm.makeWatermark(d, d, mask, out=out)
        t = time.perf_counter()
        for _ in range(a.loops):
            _, av, _ = wm.makeWatermark(d, d, mask, out=out)
        res["embed_" + name] = (time.perf_counter() - t) / a.loops
        for _ in range(2):
            wm.detectWatermark(out, mask)
        t = time.perf_counter()
        for _ in range(a.loops):
            c, _ = wm.detectWatermark(out, mask)
        res["detect_" + name] = (time.perf_counter() - t) / a.loops
        print("%s: a=%.6f corr=%.6f" % (name, av, c))
    for k, v in res.items():
        print("%-12s %8.1f us/call  %8.1f FPS" % (k, v * 1e6, 1.0 / v))
    # where the time goes: per-kernel device time (CUDA events; graphs are bypassed while timing is on)
    wm.set_option(pkg.OPT_KERNEL_TIMING, 1)
    wm.kernel_times(reset=True)
    for _ in range(50):
        for mask in (pkg.NVF, pkg.ME):
            wm.makeWatermark(d, d, mask, out=out)
            wm.detectWatermark(out, mask)
    for name, (n, ms) in wm.kernel_times(reset=True).items():
        if n:
            print("   kernel %-13s %5d launches  avg %7.2f us" % (name, n, 1e3 * ms / n))
    wm.set_option(pkg.OPT_KERNEL_TIMING, 0)
    # timeline of the Rx sweep's critical path: the CTA that finishes last (second stage + solve), from %globaltimer
    ph = []
    for _ in range(20):
        wm.detectWatermark(out, pkg.ME)
        ph.append(wm.debug(pkg.DBG_PHASES))
    import numpy as np
    ph = np.median(np.array(ph), axis=0) / 1e3
    if ph[6] == 0:
        print("   rx_sweep last CTA (us since its start): tiles %.1f | ring %.1f | elected %.1f | second-stage sums %.1f | solved %.1f" % tuple(ph[1:6]))
    if ph[6] > 0:  # the fused single-image detector ran: its own timeline (the CTA that finished the op)
        print("   k_detect1 finishing CTA (us since its start): tiles landed %.1f | sweep tiles %.1f | ring %.1f | handed over (second stage + solve elsewhere) %.1f | detector tiles %.1f | elected %.1f | done %.1f" % tuple(ph[1:8]))
    tot = sum(res.values())
    print("all four ops: %.1f us -> %.1f frames/s (single image, synchronous calls, %dx%d)" % (tot * 1e6, 1.0 / tot, a.rows, a.cols))


if __name__ == "__main__":
    main()
