"""A/B of the host-frame video path (wm_process_frames, frames_on_device = 0): padded linear upload vs 2-D repack, run sizes.
usage: python tools/e2e_video_probe.py"""
import importlib, os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import util
pkg = importlib.import_module("watermarking-gpu_b200")
rows, cols, n = 2160, 3840, 128
W = util.normal_w(rows, cols)
wm = pkg.Watermark(rows, cols, W, 3, 40.0)
base = util.natural_image(rows, cols, seed=1, integer=True)
for pad in (64, 0):
    linesize = cols + pad
    pin_in = torch.zeros((n, rows, linesize), dtype=torch.uint8).pin_memory()
    for i in range(n):
        pin_in[i, :, :cols] = torch.from_numpy(np.roll(base, (3 * i, 5 * i), (0, 1)))
    pin_out = torch.empty((n, rows, cols), dtype=torch.uint8).pin_memory()
    sc = np.zeros(2 * n, np.float32)
    vh = pkg.VideoProcessingContext(wm, rows, cols, 1, linesize=linesize, frames_on_device=False)
    ref = None
    for padded in (0, 1):
        for run in (4, 8):
            wm.set_option(pkg.OPT_PADDED_UPLOAD, padded)
            wm.set_option(pkg.OPT_HOST_RUN_FRAMES, run)
            pkg.process_frames(vh, pkg.VIDEO_EMBED_VERIFY, pin_in.data_ptr(), pin_out.data_ptr(), 0, n, sc)
            t = time.perf_counter()
            for _ in range(4):
                pkg.process_frames(vh, pkg.VIDEO_EMBED_VERIFY, pin_in.data_ptr(), pin_out.data_ptr(), 0, n, sc)
            dt = (time.perf_counter() - t) / 4
            if ref is None:
                ref = (sc.copy(), pin_out.clone())
            same = bool(np.array_equal(ref[0], sc, equal_nan=True) and torch.equal(ref[1], pin_out))
            print("linesize %d padded_upload %d run %d: %.0f frames/s, %.1f GB/s each way, identical to the first variant: %s" % (
                linesize, padded, run, n / dt, n * rows * cols / dt / 1e9, same))
