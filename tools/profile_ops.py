"""Small fixed workload for ncu: every op of the hot path, `rounds` times, on `batch` frames.
usage: python tools/profile_ops.py [--workload image1080p|image4k|video4k] [--batch 8] [--rounds 3]
Kernel order per round (image): nvf_stats, apply | sweep, me_stats, apply | sweep, detect(NVF) | sweep, detect(ME) = 9 launches;
(video, per frame): sweep, me_stats, apply, then sweep, detect = 5 launches."""
import argparse
import importlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="image1080p")
    ap.add_argument("--batch", type=int, default=8)
    ap.add_argument("--rounds", type=int, default=3)
    ap.add_argument("--no-tma", action="store_true")
    a = ap.parse_args()
    pkg = importlib.import_module("watermarking-gpu_b200")
    rows, cols, _, _, kind, dtype = bench.WORKLOADS[a.workload]
    frames, W = bench.make_inputs(rows, cols, a.batch, dtype)
    wm = pkg.Watermark(rows, cols, W, 3, 40.0)
    if a.no_tma:
        wm.set_option(pkg.OPT_USE_TMA, 0)
    L = pkg.lib()
    layout = pkg.ROW_MAJOR if kind == "video" else pkg.COL_MAJOR
    mem = frames if layout == pkg.ROW_MAJOR else np.ascontiguousarray(frames.transpose(0, 2, 1))
    din = L.wm_dev_alloc(wm._h, mem.nbytes)
    dout = L.wm_dev_alloc(wm._h, mem.nbytes)
    L.wm_dev_upload(wm._h, din, mem.ctypes.data, mem.nbytes)
    dt = pkg.U8 if dtype == "u8" else pkg.F32
    di = pkg.image_desc(din, rows, cols, layout, dt)
    do = pkg.image_desc(dout, rows, cols, layout, dt)
    npx = rows * cols
    a_h = np.zeros(a.batch, np.float32)
    c_h = np.zeros(a.batch, np.float32)
    for r in range(a.rounds):
        if kind == "video":
            v = pkg.VideoProcessingContext(wm, rows, cols, 1, linesize=cols, frames_on_device=True)
            pkg.process_frames(v, pkg.VIDEO_EMBED, din, dout, 0, a.batch, a_h)
            pkg.process_frames(v, pkg.VIDEO_DETECT, dout, None, 0, a.batch, c_h)
        else:
            for mask in (pkg.NVF, pkg.ME):
                wm.embed_batch(0, di, di, do, npx, npx, npx, a.batch, mask, a_h)
                wm.sync(0)
            for mask in (pkg.NVF, pkg.ME):
                wm.detect_batch(0, do, npx, a.batch, mask, c_h)
                wm.sync(0)
    print("ok a=%.5f corr=%.5f launches=%d" % (a_h[0], c_h[0], wm.launch_count))
    wm.close()


if __name__ == "__main__":
    main()
