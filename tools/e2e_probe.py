import importlib, sys, time, os
import numpy as np, torch
sys.path.insert(0, os.getcwd()); sys.path.insert(0, "tests")
import util
pkg = importlib.import_module("watermarking-gpu_b200")
rows, cols, n = 1080, 1920, 96
W = util.normal_w(rows, cols)
wm = pkg.Watermark(rows, cols, W, 3, 40.0)
img = util.natural_image(rows, cols, seed=1)
pin_in = torch.from_numpy(np.ascontiguousarray(np.broadcast_to(img.T, (n, cols, rows)))).pin_memory()
pin_out = [torch.empty_like(pin_in).pin_memory() for _ in range(2)]
a_e = [np.zeros(n, np.float32) for _ in range(2)]; c_e = [np.zeros(n, np.float32) for _ in range(2)]
npx = rows * cols
NS = wm.num_slots
def one_pass(chunk, masks, alt):
    for ci, o in enumerate(range(0, n, chunk)):
        nb = min(chunk, n - o)
        hin = pkg.image_desc(pin_in[o].data_ptr(), rows, cols, pkg.COL_MAJOR, pkg.F32)
        for k2, mask in enumerate(masks):
            sl = ((len(masks) * ci + k2) if alt else ci) % NS
            hout = pkg.image_desc(pin_out[k2][o].data_ptr(), rows, cols, pkg.COL_MAJOR, pkg.F32)
            wm.embed_verify_host_batch(sl, hin, hin, hout, npx, npx, npx, nb, mask, a_e[k2][o:o + nb], c_e[k2][o:o + nb])
    wm.sync(-1)
for chunk in (1, 2, 4, 8):
    for alt in (0, 1):
        one_pass(chunk, (pkg.NVF, pkg.ME), alt); torch.cuda.synchronize()
        t = time.perf_counter()
        for _ in range(4): one_pass(chunk, (pkg.NVF, pkg.ME), alt)
        torch.cuda.synchronize(); dt = (time.perf_counter() - t) / 4
        print("chunk %d alt %d: %.1f ms per pass of %d frames -> %.0f frames/s, %.1f GB/s each way" % (chunk, alt, dt * 1e3, n, n / dt, 2 * n * npx * 4 / dt / 1e9))
# host-side cost of the calls alone
t = time.perf_counter()
for ci, o in enumerate(range(0, n, 4)):
    hin = pkg.image_desc(pin_in[o].data_ptr(), rows, cols, pkg.COL_MAJOR, pkg.F32)
    hout = pkg.image_desc(pin_out[0][o].data_ptr(), rows, cols, pkg.COL_MAJOR, pkg.F32)
    wm.embed_verify_host_batch(ci % NS, hin, hin, hout, npx, npx, npx, 4, pkg.ME, a_e[0][o:o + 4], c_e[0][o:o + 4])
t1 = time.perf_counter() - t
wm.sync(-1)
print("issuing 24 calls took %.2f ms of host time (%.0f us per call)" % (t1 * 1e3, t1 / 24 * 1e6))
