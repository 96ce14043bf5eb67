"""The reference's sample application (main.cpp:62-316) as a thin CLI over libwm_b200: same settings.ini keys, same flow,
same console output.

  python tools/sample_app.py [settings.ini]

[paths] image / watermark, [options] save_watermarked_files_to_disk / execution_time_in_fps, [parameters] p / psnr /
loops_for_test are read exactly like the reference does (inih semantics: missing keys fall back to the reference's
defaults).  Image I/O is Pillow's (the reference uses ArrayFire's FreeImage loader); everything between loadImage and
saveImage runs on the device through the C ABI: rgb2gray (main.cpp:154), makeWatermark into the RGB image with the NVF and
the ME mask (main.cpp:175-193), rgb2gray of the watermarked images (main.cpp:196-197), detectWatermark (main.cpp:203-222).
[paths] video needs an ffmpeg demuxer and is outside this library: raw yuv420p files are handled by
`wm_process_frames` / csrc/videoprocessingcontext.hpp.
"""
import configparser
import importlib
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def execution_time(show_fps, seconds):  # Utilities.cpp: executionTime
    return "FPS: %.2f FPS" % (1.0 / seconds) if show_fps else "%.6f seconds" % seconds


def add_suffix(path, suffix):  # Utilities::addSuffixBeforeExtension
    base, ext = os.path.splitext(path)
    return base + suffix + ext


def main():
    ini_path = sys.argv[1] if len(sys.argv) > 1 else "settings.ini"
    ini = configparser.ConfigParser(inline_comment_prefixes=(";", "#"))
    if not ini.read(ini_path):
        print("Could not load settings.ini file")
        return 1

    def get(sec, key, default):
        return ini.get(sec, key, fallback=default).strip()

    p = int(get("parameters", "p", "-1"))
    psnr = float(get("parameters", "psnr", "-1"))
    if p != 3:  # main.cpp:89
        print("For now, only p=3 is allowed")
        return 1
    if psnr <= 0:
        print("PSNR must be a positive number")
        return 1
    if get("paths", "video", ""):
        print("video input needs ffmpeg demux/decode, which is outside this library (see DESIGN.md §7)")
        return 1
    image_file = get("paths", "image", "NO_IMAGE")
    show_fps = get("options", "execution_time_in_fps", "false").lower() == "true"
    loops = int(get("parameters", "loops_for_test", "5"))
    loops = 5 if loops <= 0 else loops
    print("Each test will be executed %d times. Average time will be shown below" % loops)

    pkg = importlib.import_module("watermarking-gpu_b200")
    from PIL import Image
    t = time.perf_counter()
    rgb = np.ascontiguousarray(np.asarray(Image.open(image_file).convert("RGB"), np.float32).transpose(2, 0, 1))
    rows, cols = rgb.shape[1:]
    if cols < 64 or rows < 64:
        print("Image dimensions too low")
        return 1
    # Watermark(rows, cols, path, p, psnr): wrong p / W size raise like the reference (Watermark.cpp:24-25,65-71)
    wm = pkg.Watermark(rows, cols, get("paths", "watermark", ""), p, psnr)
    d_rgb = pkg.DeviceArray.from_numpy(wm, rgb, pkg.COL_MAJOR)
    d_gray = wm.rgb2gray(d_rgb)
    wm.sync(-1)
    print("Time to load and transfer RGB image from disk to VRAM: %.6f\n" % (time.perf_counter() - t))

    outs = {}
    for name, mask in (("NVF", pkg.NVF), ("ME", pkg.ME)):  # warm-up (main.cpp:169-170)
        outs[name] = pkg.DeviceArray(wm, rows, cols, pkg.COL_MAJOR, pkg.F32, 3)
        wm.makeWatermark(d_gray, d_rgb, mask, out=outs[name])
    for name, mask in (("NVF", pkg.NVF), ("ME", pkg.ME)):
        secs, a = 0.0, 0.0
        for _ in range(loops):
            t = time.perf_counter()
            _, a, _ = wm.makeWatermark(d_gray, d_rgb, mask, out=outs[name])
            secs += time.perf_counter() - t
        print("Watermark strength (parameter a): %s\nCalculation of %s mask with %d rows and %d columns and parameters:\np = %d  PSNR(dB) = %s\n%s\n"
              % (a, name, rows, cols, p, psnr, execution_time(show_fps, secs / loops)))
    grays = {name: wm.rgb2gray(outs[name]) for name in outs}
    corr = {}
    for name, mask in (("NVF", pkg.NVF), ("ME", pkg.ME)):
        wm.detectWatermark(grays[name], mask)  # warm-up (main.cpp:199-200)
    for name, mask in (("NVF", pkg.NVF), ("ME", pkg.ME)):
        secs = 0.0
        for _ in range(loops):
            t = time.perf_counter()
            corr[name], _ = wm.detectWatermark(grays[name], mask)
            secs += time.perf_counter() - t
        print("Calculation of the watermark correlation (%s) of an image with %d rows and %d columns and parameters:\np = %d  PSNR(dB) = %s\n%s\n"
              % (name, rows, cols, p, psnr, execution_time(show_fps, secs / loops)))
    print("Correlation [NVF]: %.16f" % corr["NVF"])
    print("Correlation [ME]: %.16f" % corr["ME"])
    if get("options", "save_watermarked_files_to_disk", "false").lower() == "true":
        print("\nSaving watermarked files to disk...")
        for name in outs:  # .as(u8): truncation (main.cpp:236-238)
            img = outs[name].numpy().astype(np.uint8).transpose(1, 2, 0)
            Image.fromarray(img).save(add_suffix(image_file, "_W_" + name))
        print("Successully saved to disk")
    wm.close()
    return 0


if __name__ == "__main__":
    sys.exit(main())
