"""GPU parity tests: the CUDA path (through the C ABI) against the CPU oracle on the same seeded inputs.

Tolerances are north_star's: masks / prediction error 1e-4 relative (scale-relative),
`a` and correlation 1e-3 relative, 8-bit pixels +-1 LSB.  Integer work (Rx/rx on integer-valued pixels, the
u8 video frames) is checked bit-exactly.
"""
import os

import numpy as np
import pytest

import util

pytestmark = pytest.mark.gpu

REPORT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out", "parity_report.txt")


def report(msg):
    os.makedirs(os.path.dirname(REPORT), exist_ok=True)
    with open(REPORT, "a") as f:
        f.write(msg + "\n")
    print(msg)


LAYOUTS = [0, 1]  # COL_MAJOR (ArrayFire), ROW_MAJOR (video)
SIZES = [(64, 64), (67, 131), (200, 96), (130, 260), (512, 512)]


def _mk(wmb, rows, cols, W, psnr=40.0):
    return wmb.Watermark(rows, cols, W, 3, psnr)


# ---------------------------------------------------------------------------------------------------
# stage (i): Rx / rx
# ---------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("layout", LAYOUTS)
@pytest.mark.parametrize("rows,cols", SIZES)
def test_rx_bit_exact_integer_pixels(wmb, oracle, rows, cols, layout):
    """u8 frames: every fp16-rounded product is an integer, every partial sum is exact => Rx, rx bit-exact."""
    img = util.natural_image(rows, cols, seed=rows * 7 + cols, integer=True)
    W = util.normal_w(rows, cols)
    wm = _mk(wmb, rows, cols, W)
    for fp16, o in ((1, oracle.FAITHFUL), (0, oracle.EXACT)):
        wm.set_option(wmb.OPT_FP16_PRODUCTS, fp16)
        for dt in (np.uint8, np.float32):
            d = wmb.DeviceArray.from_numpy(wm, img.astype(dt), layout)
            wm.detectWatermark(d, wmb.ME)
            Rx, rx = wm.debug(wmb.DBG_RX), wm.debug(wmb.DBG_RXVEC)
            oRx, orx = oracle.rx(img.astype(np.float32), o)
            report("rx_int %dx%d layout=%d fp16=%d dt=%s maxdiff Rx=%g rx=%g" % (
                rows, cols, layout, fp16, dt.__name__, np.abs(Rx - oRx).max(), np.abs(rx - orx).max()))
            assert np.array_equal(Rx, oRx)
            assert np.array_equal(rx, orx)
    wm.close()


@pytest.mark.parametrize("layout", LAYOUTS)
def test_rx_f32_and_coeffs(wmb, oracle, layout):
    img = util.load_512_gray(oracle)
    W = util.load_w512()
    wm = _mk(wmb, 512, 512, util.W512_PATH)
    d = wmb.DeviceArray.from_numpy(wm, img, layout)
    wm.detectWatermark(d, wmb.ME)
    Rx, rx, c = wm.debug(wmb.DBG_RX), wm.debug(wmb.DBG_RXVEC), wm.debug(wmb.DBG_COEFFS)
    oRx, orx = oracle.rx(img, oracle.FAITHFUL)
    r1, r2 = util.rel(Rx, oRx), util.rel(rx, orx)
    report("rx_f32 512 layout=%d rel Rx=%.3g rx=%.3g" % (layout, r1, r2))
    assert r1 <= 1e-6 and r2 <= 1e-6  # SURVEY H1 stage (i)
    # stage (ii): the solve, against the oracle's solve of the SAME system
    st, oc = oracle.solve8(Rx, rx)
    assert st == 0
    res = np.linalg.norm(Rx @ c.astype(np.float64) - rx) / np.linalg.norm(rx)
    report("solve 512 layout=%d max|c-oc|=%.3g residual=%.3g" % (layout, np.abs(c - oc.astype(np.float32)).max(), res))
    assert np.array_equal(c, oc.astype(np.float32))
    # and end to end against the oracle's own coefficients
    pe = oracle.pred_error_mask(img, oracle.FAITHFUL)
    rc = util.rel(c, pe["coef"])
    report("coef 512 layout=%d rel vs oracle=%.3g" % (layout, rc))
    assert rc <= 1e-4
    wm.close()
    del W


# ---------------------------------------------------------------------------------------------------
# stage (iii): prediction error and masks with the oracle's coefficients injected
# ---------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("layout", LAYOUTS)
@pytest.mark.parametrize("rows,cols", [(64, 64), (67, 131), (512, 512)])
def test_errseq_and_nvf_planes(wmb, oracle, rows, cols, layout):
    img = util.natural_image(rows, cols, seed=3)
    W = util.normal_w(rows, cols)
    wm = _mk(wmb, rows, cols, W)
    d = wmb.DeviceArray.from_numpy(wm, img, layout)
    pe = oracle.pred_error_mask(img, oracle.FAITHFUL)
    wm.debug_set_coeffs(pe["coef"])
    e = wm.debug_plane(d, wmb.DBG_ERRSEQ)
    wm.debug_set_coeffs(None)
    nv = wm.debug_plane(d, wmb.DBG_MASK_NVF)
    onv = oracle.nvf(img, oracle.FAITHFUL)
    re_, rn = util.rel(e, pe["e"]), util.rel(nv, onv)
    report("planes %dx%d layout=%d e: rel=%.3g exact=%s | nvf: rel=%.3g exact=%s" % (
        rows, cols, layout, re_, np.array_equal(e, pe["e"]), rn, np.array_equal(nv, onv)))
    assert re_ <= 1e-4 and rn <= 1e-4
    # same operation order as the reference kernels => in fact bit-identical
    assert np.array_equal(e, pe["e"])
    assert np.array_equal(nv, onv)
    # free-running (own Rx sweep + solve): still within tolerance
    e2 = wm.debug_plane(d, wmb.DBG_ERRSEQ)
    r2 = util.rel(e2, pe["e"])
    report("planes %dx%d layout=%d e (own coefficients): rel=%.3g" % (rows, cols, layout, r2))
    assert r2 <= 1e-4
    wm.close()


# ---------------------------------------------------------------------------------------------------
# stage (iv): end to end — makeWatermark / detectWatermark
# ---------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("layout", LAYOUTS)
@pytest.mark.parametrize("mask", [0, 1])
def test_embed_detect_512_fixture(wmb, oracle, mask, layout):
    """BASELINE config 1: 512.png + w_512.dat, p=3, psnr=40, gray in / gray out."""
    img = util.load_512_gray(oracle)
    W = util.load_w512()
    wm = _mk(wmb, 512, 512, util.W512_PATH)
    d = wmb.DeviceArray.from_numpy(wm, img, layout)
    out, a, st = wm.makeWatermark(d, d, mask)
    o = oracle.embed(img, W, 40.0, mask)
    assert st == 0 and o["status"] == 0
    got = out.numpy()
    ra = abs(a - o["a"]) / abs(o["a"])
    dpx = np.abs(got - o["out"]).max()
    report("embed512 mask=%d layout=%d a=%.7g oracle=%.7g rel=%.3g max|dpix|=%.3g" % (mask, layout, a, o["a"], ra, dpx))
    assert ra <= 1e-3
    assert dpx <= 1e-4 * 255
    # 8-bit pixels +-1 LSB
    q = np.abs(got.astype(np.uint8).astype(int) - o["out"].astype(np.uint8).astype(int)).max()
    assert q <= 1
    # detection on the watermarked image, and on the clean one
    for name, z in (("marked", o["out"]), ("clean", img)):
        dz = wmb.DeviceArray.from_numpy(wm, z, layout)
        corr, st = wm.detectWatermark(dz, mask)
        od = oracle.detect(z, W, mask)
        rcorr = abs(corr - od["corr"]) / max(abs(od["corr"]), 1e-3)
        report("detect512 %s mask=%d layout=%d corr=%.7g oracle=%.7g rel=%.3g" % (name, mask, layout, corr, od["corr"], rcorr))
        assert st == 0
        assert rcorr <= 1e-3
    wm.close()


@pytest.mark.parametrize("layout", LAYOUTS)
def test_embed_rgb_base(wmb, oracle, layout):
    """testForImage flow: gray input, RGB output image (main.cpp:178,190; Watermark.cpp:171 broadcast)."""
    rgb = util.load_512_rgb()
    gray = oracle.rgb2gray(rgb)
    W = util.load_w512()
    wm = _mk(wmb, 512, 512, W)
    dg = wmb.DeviceArray.from_numpy(wm, gray, layout)
    drgb = wmb.DeviceArray.from_numpy(wm, rgb, layout)
    for mask in (wmb.NVF, wmb.ME):
        out, a, st = wm.makeWatermark(dg, drgb, mask)
        o = oracle.embed(gray, W, 40.0, mask, base=rgb)
        got = out.numpy()
        report("embed_rgb mask=%d layout=%d a rel=%.3g max|dpix|=%.3g" % (
            mask, layout, abs(a - o["a"]) / o["a"], np.abs(got - o["out"]).max()))
        assert got.shape == (3, 512, 512)
        assert abs(a - o["a"]) / o["a"] <= 1e-3
        assert np.abs(got - o["out"]).max() <= 1e-4 * 255
    wm.close()


@pytest.mark.parametrize("rows,cols", [(64, 64), (67, 131), (70, 94), (200, 96), (130, 260), (96, 1030)])
def test_odd_shapes_all_ops(wmb, oracle, rows, cols):
    """Edge shapes of the reference's samples in miniature: not divisible by 4 / 16 / 64 / the tile."""
    img = util.natural_image(rows, cols, seed=rows + cols)
    W = util.normal_w(rows, cols)
    wm = _mk(wmb, rows, cols, W)
    for layout in LAYOUTS:
        d = wmb.DeviceArray.from_numpy(wm, img, layout)
        for mask in (wmb.ME, wmb.NVF):
            out, a, st = wm.makeWatermark(d, d, mask)
            o = oracle.embed(img, W, 40.0, mask)
            got = out.numpy()
            dz = wmb.DeviceArray.from_numpy(wm, o["out"], layout)
            corr, st2 = wm.detectWatermark(dz, mask)
            od = oracle.detect(o["out"], W, mask)
            report("odd %dx%d layout=%d mask=%d a rel=%.3g dpix=%.3g corr rel=%.3g" % (
                rows, cols, layout, mask, abs(a - o["a"]) / o["a"], np.abs(got - o["out"]).max(),
                abs(corr - od["corr"]) / abs(od["corr"])))
            assert st == 0 and st2 == 0
            assert abs(a - o["a"]) / o["a"] <= 1e-3
            assert np.abs(got - o["out"]).max() <= 1e-4 * 255
            assert abs(corr - od["corr"]) / abs(od["corr"]) <= 1e-3
    wm.close()


def test_strided_input_and_output(wmb, oracle):
    rows, cols = 72, 100
    img = util.natural_image(rows, cols, seed=5)
    W = util.normal_w(rows, cols)
    wm = _mk(wmb, rows, cols, W)
    for layout in LAYOUTS:
        P = rows if layout == wmb.COL_MAJOR else cols
        d = wmb.DeviceArray.from_numpy(wm, img, layout, ld=P + 12)
        dout = wmb.DeviceArray(wm, rows, cols, layout, wmb.F32, 1, ld=P + 7)  # unaligned ld: scalar store path
        out, a, st = wm.makeWatermark(d, d, wmb.ME, out=dout)
        o = oracle.embed(img, W, 40.0, wmb.ME)
        assert abs(a - o["a"]) / o["a"] <= 1e-3
        assert np.abs(out.numpy() - o["out"]).max() <= 1e-4 * 255
        corr, _ = wm.detectWatermark(d, wmb.NVF)
        od = oracle.detect(img, W, wmb.NVF)
        assert abs(corr - od["corr"]) <= 1e-3 * max(abs(od["corr"]), 1e-2)
    wm.close()


# ---------------------------------------------------------------------------------------------------
# u8 video frames (main.cpp:343-410): bit-exact pipeline
# ---------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("rows,cols,linesize", [(64, 64, 64), (120, 200, 224), (270, 480, 512)])
def test_video_frames_u8(wmb, oracle, rows, cols, linesize):
    n = 7
    interval = 2
    W = util.normal_w(rows, cols)
    wm = _mk(wmb, rows, cols, W)
    frames = np.zeros((n, rows, linesize), np.uint8)
    for i in range(n):
        frames[i, :, :cols] = util.natural_image(rows, cols, seed=100 + i, integer=True)
        frames[i, :, cols:] = 255 - (i * 13) % 200  # row padding must be ignored
    ctx = wmb.VideoProcessingContext(wm, rows, cols, interval, linesize=linesize, frames_on_device=False)
    out = np.zeros((n, rows, cols), np.uint8)
    a = np.zeros(n, np.float32)
    first = 10  # global index offset (sharding): gating uses first + i
    got = wmb.process_frames(ctx, wmb.VIDEO_EMBED, frames.ctypes.data, out.ctypes.data, first, n, a)
    assert got == n
    corr = np.zeros(n, np.float32)
    ctx2 = wmb.VideoProcessingContext(wm, rows, cols, interval, linesize=cols, frames_on_device=False)
    wmb.process_frames(ctx2, wmb.VIDEO_DETECT, out.ctypes.data, None, first, n, corr)
    nexact = 0
    for i in range(n):
        if (first + i) % interval:
            assert np.array_equal(out[i], frames[i, :, :cols])
            assert np.isnan(a[i]) and np.isnan(corr[i])
            continue
        st, oo, oa = oracle.embed_frame_u8(frames[i], W, 40.0, oracle.ME, width=cols)
        st2, oc = oracle.detect_frame_u8(oo, W, oracle.ME)
        dmax = np.abs(out[i].astype(int) - oo.astype(int)).max()
        nexact += int(np.array_equal(out[i], oo))
        report("video %dx%d ls=%d frame %d: a=%.7g/%.7g dpix=%d corr=%.7g/%.7g" % (
            rows, cols, linesize, i, a[i], oa, dmax, corr[i], oc))
        assert dmax <= 1
        assert abs(a[i] - oa) / oa <= 1e-3
        assert abs(corr[i] - oc) / abs(oc) <= 1e-3
    report("video %dx%d: %d gated frames bit-identical to the oracle" % (rows, cols, nexact))
    wm.close()


def test_video_device_frames_and_sharding(wmb, oracle):
    """Frames resident on the device, strided reads (no repack pass); two shards == one pass."""
    rows, cols, ls, n = 96, 160, 192, 6
    W = util.normal_w(rows, cols)
    wm = _mk(wmb, rows, cols, W)
    frames = np.zeros((n, rows, ls), np.uint8)
    for i in range(n):
        frames[i, :, :cols] = util.natural_image(rows, cols, seed=200 + i, integer=True)
    L = wmb.lib()
    dfr = L.wm_dev_alloc(wm._h, frames.nbytes)
    dout = L.wm_dev_alloc(wm._h, n * rows * cols)
    L.wm_dev_upload(wm._h, dfr, frames.ctypes.data, frames.nbytes)
    ctx = wmb.VideoProcessingContext(wm, rows, cols, 1, linesize=ls, frames_on_device=True)
    a_all = np.zeros(n, np.float32)
    wmb.process_frames(ctx, wmb.VIDEO_EMBED, dfr, dout, 0, n, a_all)
    out_all = np.zeros((n, rows, cols), np.uint8)
    L.wm_dev_download(wm._h, out_all.ctypes.data, dout, out_all.nbytes)
    # shards
    a_sh = np.zeros(n, np.float32)
    for rank in range(2):
        first, cnt = wmb.shard_frames(n, rank, 2)
        wmb.process_frames(ctx, wmb.VIDEO_EMBED, dfr + first * rows * ls, dout + first * rows * cols, first, cnt,
                           a_sh[first:first + cnt])
    out_sh = np.zeros_like(out_all)
    L.wm_dev_download(wm._h, out_sh.ctypes.data, dout, out_sh.nbytes)
    assert np.array_equal(out_all, out_sh) and np.array_equal(a_all, a_sh)
    for i in (0, n - 1):
        st, oo, oa = oracle.embed_frame_u8(frames[i], W, 40.0, oracle.ME, width=cols)
        assert np.abs(out_all[i].astype(int) - oo.astype(int)).max() <= 1
        assert abs(a_all[i] - oa) / oa <= 1e-3
    L.wm_dev_free(wm._h, dfr)
    L.wm_dev_free(wm._h, dout)
    wm.close()


# ---------------------------------------------------------------------------------------------------
# batched launch (BASELINE config 5 in miniature) and host-buffer API
# ---------------------------------------------------------------------------------------------------
def test_batched_equals_single(wmb, oracle):
    rows = cols = 64
    B = 9
    W = util.normal_w(rows, cols)
    wm = _mk(wmb, rows, cols, W)
    imgs = np.stack([util.natural_image(rows, cols, seed=300 + b) for b in range(B)])
    imgs[4] = 17.0  # a constant image in the middle of the batch: singular system
    L = wmb.lib()
    din = L.wm_dev_alloc(wm._h, imgs.nbytes)
    dout = L.wm_dev_alloc(wm._h, imgs.nbytes)
    L.wm_dev_upload(wm._h, din, imgs.ctypes.data, imgs.nbytes)
    for mask in (wmb.ME, wmb.NVF):
        a = np.full(B, np.nan, np.float32)
        st = np.zeros(B, np.int32)
        di = wmb.image_desc(din, rows, cols, wmb.ROW_MAJOR, wmb.F32)
        do = wmb.image_desc(dout, rows, cols, wmb.ROW_MAJOR, wmb.F32)
        wm.embed_batch(1, di, di, do, rows * cols, rows * cols, rows * cols, B, mask, a, st)
        wm.sync(1)
        outs = np.zeros_like(imgs)
        L.wm_dev_download(wm._h, outs.ctypes.data, dout, outs.nbytes)
        corr = np.zeros(B, np.float32)
        st2 = np.zeros(B, np.int32)
        wm.detect_batch(2, do, rows * cols, B, mask, corr, st2)
        wm.sync(2)
        for b in range(B):
            o = oracle.embed(imgs[b], W, 40.0, mask)
            if b == 4:
                assert st[b] != 0 and np.array_equal(outs[b], imgs[b])
                continue
            od = oracle.detect(o["out"], W, mask)
            report("batch mask=%d b=%d a rel=%.3g dpix=%.3g corr rel=%.3g" % (
                mask, b, abs(a[b] - o["a"]) / o["a"], np.abs(outs[b] - o["out"]).max(),
                abs(corr[b] - od["corr"]) / abs(od["corr"])))
            assert st[b] == 0 and st2[b] == 0
            assert abs(a[b] - o["a"]) / o["a"] <= 1e-3
            assert np.abs(outs[b] - o["out"]).max() <= 1e-4 * 255
            assert abs(corr[b] - od["corr"]) / abs(od["corr"]) <= 1e-3
    L.wm_dev_free(wm._h, din)
    L.wm_dev_free(wm._h, dout)
    wm.close()


def test_host_buffer_api(wmb, oracle):
    rows, cols = 128, 192
    img = util.natural_image(rows, cols, seed=11)
    W = util.normal_w(rows, cols)
    wm = _mk(wmb, rows, cols, W)
    out = np.zeros_like(img)
    a, st = wm.make_watermark_host(img, img, out, wmb.ME, layout=wmb.ROW_MAJOR)
    o = oracle.embed(img, W, 40.0, wmb.ME)
    assert st == 0 and abs(a - o["a"]) / o["a"] <= 1e-3
    assert np.abs(out - o["out"]).max() <= 1e-4 * 255
    corr, st = wm.detect_watermark_host(out, wmb.ME, layout=wmb.ROW_MAJOR)
    od = oracle.detect(o["out"], W, wmb.ME)
    assert abs(corr - od["corr"]) / abs(od["corr"]) <= 1e-3
    wm.close()


# ---------------------------------------------------------------------------------------------------
# error behaviour of the reference (Watermark.cpp:24-25,65-71,164-165,246-247)
# ---------------------------------------------------------------------------------------------------
def test_singular_system(wmb, oracle):
    rows = cols = 64
    img = np.full((rows, cols), 100.0, np.float32)
    W = util.normal_w(rows, cols)
    wm = _mk(wmb, rows, cols, W)
    d = wmb.DeviceArray.from_numpy(wm, img, wmb.COL_MAJOR)
    out, a, st = wm.makeWatermark(d, d, wmb.ME)
    assert st == wmb.SINGULAR and np.isnan(a)          # `a` untouched, image returned unchanged
    assert np.array_equal(out.numpy(), img)
    corr, st = wm.detectWatermark(d, wmb.ME)
    assert st == wmb.SINGULAR and corr == 0.0
    corr, st = wm.detectWatermark(d, wmb.NVF)
    assert st == wmb.SINGULAR and corr == 0.0
    o = oracle.embed(img, W, 40.0, wmb.ME)
    assert o["status"] == 1
    # NVF of a constant image is identically 0: reference gives a = inf / NaN pixels; we return the base
    out, a, st = wm.makeWatermark(d, d, wmb.NVF)
    assert st == wmb.ZERO_MASK and np.isinf(a) and np.array_equal(out.numpy(), img)
    wm.close()


def test_argument_errors(wmb, tmp_path):
    W = util.normal_w(64, 64)
    with pytest.raises(wmb.WatermarkError) as e:
        wmb.Watermark(64, 64, W, 4, 40.0)
    assert e.value.code == -1 and "Wrong p parameter" in str(e.value)
    with pytest.raises(wmb.WatermarkError) as e:
        wmb.Watermark(64, 64, str(tmp_path / "missing.dat"), 3, 40.0)
    assert e.value.code == -2
    bad = tmp_path / "short.dat"
    W[:10].tofile(bad)
    with pytest.raises(wmb.WatermarkError) as e:
        wmb.Watermark(64, 64, str(bad), 3, 40.0)
    assert e.value.code == -3 and "W file total elements != image dimensions" in str(e.value)
    wm = wmb.Watermark(64, 64, W, 3, 40.0)
    d = wmb.DeviceArray(wm, 64, 80)
    with pytest.raises(wmb.WatermarkError) as e:
        wm.detectWatermark(d, wmb.ME)
    assert e.value.code == -4
    # reinitialize to the new size, clone shares W
    W2 = util.normal_w(64, 80)
    wm.reinitialize(W2, 64, 80)
    img = util.natural_image(64, 80, seed=1)
    d.upload(img)
    c1, _ = wm.detectWatermark(d, wmb.ME)
    wm2 = wm.clone()
    c2, _ = wm2.detectWatermark(d, wmb.ME)
    assert c1 == c2
    wm2.close()
    wm.close()


# ---------------------------------------------------------------------------------------------------
# full-size, size-independent properties (BASELINE configs 2-4)
# ---------------------------------------------------------------------------------------------------
# 1078x1918 and 2160x3872 are the reference's own odd-size samples (not divisible by 16 / by 64)
@pytest.mark.parametrize("rows,cols", [(1080, 1920), (1078, 1918), (2160, 3840), (2160, 3872)])
def test_full_size_properties(wmb, oracle, rows, cols):
    img = util.natural_image(rows, cols, seed=42)
    W = util.normal_w(rows, cols)
    wm = _mk(wmb, rows, cols, W)
    sf = wm.strength_factor
    res = {}
    for layout in LAYOUTS:
        d = wmb.DeviceArray.from_numpy(wm, img, layout)
        for mask in (wmb.ME, wmb.NVF):
            out, a, st = wm.makeWatermark(d, d, mask)
            got = out.numpy()
            assert st == 0
            # determinism: a second run is bit-identical
            out2, a2, _ = wm.makeWatermark(d, d, mask)
            assert a2 == a and np.array_equal(out2.numpy(), got)
            # PSNR property: MSE == strength^2 before clamping (Watermark.cpp:170); clamping only lowers it
            mse = float(np.mean((got.astype(np.float64) - img) ** 2))
            report("full %dx%d layout=%d mask=%d a=%.7g mse/sf^2=%.6f" % (rows, cols, layout, mask, a, mse / sf ** 2))
            assert 0.9 <= mse / sf ** 2 <= 1.0 + 1e-4
            dz = wmb.DeviceArray.from_numpy(wm, got, layout)
            cm, _ = wm.detectWatermark(dz, mask)
            cc, _ = wm.detectWatermark(d, mask)
            report("full %dx%d layout=%d mask=%d corr marked=%.6f clean=%.6f" % (rows, cols, layout, mask, cm, cc))
            assert cm > 0.2 and abs(cc) < 0.05
            res[(layout, mask)] = (a, got, cm)
    # layout equivalence: the transposed code path gives the same answer (not bitwise: summation order differs)
    for mask in (wmb.ME, wmb.NVF):
        a0, g0, c0 = res[(0, mask)]
        a1, g1, c1 = res[(1, mask)]
        assert abs(a0 - a1) / a1 <= 1e-5 and abs(c0 - c1) <= 1e-5
        assert np.abs(g0 - g1).max() <= 5e-3
    # spot-check against the oracle at full size for the ME mask (CPU work: a few seconds)
    if rows == 1080:
        o = oracle.embed(img, W, 40.0, wmb.ME)
        a1, g1, _ = res[(1, wmb.ME)]
        report("full 1080p ME vs oracle: a rel=%.3g dpix=%.3g" % (abs(a1 - o["a"]) / o["a"], np.abs(g1 - o["out"]).max()))
        assert abs(a1 - o["a"]) / o["a"] <= 1e-3
        assert np.abs(g1 - o["out"]).max() <= 1e-4 * 255
    wm.close()


# ---------------------------------------------------------------------------------------------------
# TMA tile pipeline vs the plain cooperative loader: same arithmetic, so bit-identical results
# ---------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("rows,cols", [(64, 64), (128, 32), (200, 96), (132, 260), (512, 512), (1080, 1920)])
def test_tma_and_plain_paths_agree(wmb, oracle, rows, cols):
    img = util.natural_image(rows, cols, seed=rows + 3 * cols)
    W = util.normal_w(rows, cols)
    wm = _mk(wmb, rows, cols, W)
    res = {}
    for tma in (1, 0):
        wm.set_option(wmb.OPT_USE_TMA, tma)
        for layout in LAYOUTS:
            d = wmb.DeviceArray.from_numpy(wm, img, layout)
            for mask in (wmb.ME, wmb.NVF):
                out, a, st = wm.makeWatermark(d, d, mask)
                got = out.numpy()
                dz = wmb.DeviceArray.from_numpy(wm, got, layout)
                corr, st2 = wm.detectWatermark(dz, mask)
                Rx = wm.debug(wmb.DBG_RX)
                res[(tma, layout, mask)] = (a, got, corr, Rx, st, st2)
    for layout in LAYOUTS:
        for mask in (wmb.ME, wmb.NVF):
            a1, g1, c1, R1, s1, t1 = res[(1, layout, mask)]
            a0, g0, c0, R0, s0, t0 = res[(0, layout, mask)]
            report("tma_vs_plain %dx%d layout=%d mask=%d a %g/%g corr %g/%g dpix=%g" % (
                rows, cols, layout, mask, a1, a0, c1, c0, np.abs(g1 - g0).max()))
            assert s1 == 0 and t1 == 0 and s0 == 0 and t0 == 0
            assert a1 == a0 and c1 == c0 and np.array_equal(g1, g0) and np.array_equal(R1, R0)
    if rows <= 512:
        o = oracle.embed(img, W, 40.0, wmb.ME)
        assert abs(res[(1, 1, wmb.ME)][0] - o["a"]) / o["a"] <= 1e-3
        assert np.abs(res[(1, 1, wmb.ME)][1] - o["out"]).max() <= 1e-4 * 255
    wm.close()


# ---------------------------------------------------------------------------------------------------
# HMMA accumulation of the fp16-rounded products (WM_OPT_MMA_ACCUM) vs the FHADD chain
# ---------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("rows,cols", [(64, 64), (67, 131), (200, 96), (512, 512), (1080, 1920), (2160, 3840)])
def test_mma_and_fhadd_accumulation_agree(wmb, oracle, rows, cols):
    """Integer-valued pixels: every rounded product and every partial sum is an integer below 2^24, so the tensor-pipe sums
    must give the same bits as the scalar chain (and, at oracle-sized shapes, as the oracle).  Real-valued pixels: f32 partial
    sums in a different association, Rx within 1e-6 and coefficients / results within north_star's tolerances."""
    W = util.normal_w(rows, cols)
    wm = _mk(wmb, rows, cols, W)
    ii = util.natural_image(rows, cols, seed=11 + rows, integer=True)
    for layout in LAYOUTS:
        for dt in (np.uint8, np.float32):
            got = {}
            for mma in (1, 0):
                wm.set_option(wmb.OPT_MMA_ACCUM, mma)
                d = wmb.DeviceArray.from_numpy(wm, ii.astype(dt), layout)
                corr, st = wm.detectWatermark(d, wmb.ME)
                got[mma] = (wm.debug(wmb.DBG_RX), wm.debug(wmb.DBG_RXVEC), wm.debug(wmb.DBG_COEFFS), corr, st)
            assert got[1][4] == 0 and got[0][4] == 0
            assert np.array_equal(got[1][0], got[0][0]) and np.array_equal(got[1][1], got[0][1])
            assert np.array_equal(got[1][2], got[0][2]) and got[1][3] == got[0][3]
            if rows <= 512:
                oRx, orx = oracle.rx(ii.astype(np.float32), oracle.FAITHFUL)
                assert np.array_equal(got[1][0], oRx) and np.array_equal(got[1][1], orx)
    img = util.natural_image(rows, cols, seed=5 + cols)
    for layout in LAYOUTS:
        got = {}
        for mma in (1, 0):
            wm.set_option(wmb.OPT_MMA_ACCUM, mma)
            d = wmb.DeviceArray.from_numpy(wm, img, layout)
            out, a, st = wm.makeWatermark(d, d, wmb.ME)
            z = out.numpy()
            corr, st2 = wm.detectWatermark(wmb.DeviceArray.from_numpy(wm, z, layout), wmb.ME)
            got[mma] = (wm.debug(wmb.DBG_RX), wm.debug(wmb.DBG_COEFFS), a, z, corr)
            assert st == 0 and st2 == 0
        rR = util.rel(got[1][0], got[0][0])
        rc = util.rel(got[1][1], got[0][1])
        ra = abs(got[1][2] - got[0][2]) / abs(got[0][2])
        rcorr = abs(got[1][4] - got[0][4]) / abs(got[0][4])
        dpx = np.abs(got[1][3] - got[0][3]).max()
        report("mma_vs_fhadd %dx%d layout=%d rel Rx=%.3g coef=%.3g a=%.3g corr=%.3g dpix=%.3g" % (rows, cols, layout, rR, rc, ra, rcorr, dpx))
        assert rR <= 1e-6 and ra <= 1e-4 and rcorr <= 1e-4 and dpx <= 1e-4 * 255
        if rows <= 512:
            # the debug Rx is that of the last sweep, i.e. of the watermarked image the detector was given
            r1 = util.rel(got[1][0], oracle.rx(got[1][3], oracle.FAITHFUL)[0])
            r0 = util.rel(got[0][0], oracle.rx(got[0][3], oracle.FAITHFUL)[0])
            report("mma_vs_fhadd %dx%d layout=%d Rx vs oracle: hmma %.3g fhadd %.3g" % (rows, cols, layout, r1, r0))
            assert r1 <= 1e-6
    wm.close()


@pytest.mark.parametrize("rows,cols,ls", [(64, 64, 64), (96, 160, 192), (270, 480, 480), (130, 264, 272)])
def test_u8_tma_and_plain_paths_agree(wmb, oracle, rows, cols, ls):
    """u8 frames resident on the device: TMA byte tiles + conversion pass vs the register-prefetched plain loader, on the same CTA shape
    (WM_OPT_NARROW_U8 = 0: the 128-thread stats / apply kernels of the TMA path group the f32 partial sums of a tile differently, which
    moves `a` in its last bit — that pair is compared in test_narrow_u8_kernels_equal_wide_ones)."""
    n = 5
    W = util.normal_w(rows, cols)
    wm = _mk(wmb, rows, cols, W)
    wm.set_option(wmb.OPT_NARROW_U8, 0)
    frames = np.zeros((n, rows, ls), np.uint8)
    for i in range(n):
        frames[i, :, :cols] = util.natural_image(rows, cols, seed=500 + i, integer=True)
        frames[i, :, cols:] = 9
    L = wmb.lib()
    dfr = L.wm_dev_alloc(wm._h, frames.nbytes)
    dout = L.wm_dev_alloc(wm._h, n * rows * cols)
    L.wm_dev_upload(wm._h, dfr, frames.ctypes.data, frames.nbytes)
    ctx = wmb.VideoProcessingContext(wm, rows, cols, 1, linesize=ls, frames_on_device=True)
    ctx_o = wmb.VideoProcessingContext(wm, rows, cols, 1, linesize=cols, frames_on_device=True)
    res = {}
    for tma in (1, 0):
        wm.set_option(wmb.OPT_USE_TMA, tma)
        a = np.zeros(n, np.float32)
        c = np.zeros(n, np.float32)
        wmb.process_frames(ctx, wmb.VIDEO_EMBED, dfr, dout, 0, n, a)
        out = np.zeros((n, rows, cols), np.uint8)
        L.wm_dev_download(wm._h, out.ctypes.data, dout, out.nbytes)
        wmb.process_frames(ctx_o, wmb.VIDEO_DETECT, dout, None, 0, n, c)
        res[tma] = (a, out, c)
    report("u8 tma_vs_plain %dx%d ls=%d a=%s corr=%s" % (rows, cols, ls, res[1][0][:2], res[1][2][:2]))
    assert np.array_equal(res[1][0], res[0][0]) and np.array_equal(res[1][1], res[0][1]) and np.array_equal(res[1][2], res[0][2])
    st, oo, oa = oracle.embed_frame_u8(frames[2], W, 40.0, oracle.ME, width=cols)
    st2, oc = oracle.detect_frame_u8(oo, W, oracle.ME)
    assert np.abs(res[1][1][2].astype(int) - oo.astype(int)).max() <= 1
    assert abs(res[1][0][2] - oa) / oa <= 1e-3 and abs(res[1][2][2] - oc) / abs(oc) <= 1e-3
    L.wm_dev_free(wm._h, dfr)
    L.wm_dev_free(wm._h, dout)
    wm.close()


def test_many_small_images_one_launch(wmb, oracle):
    """BASELINE config 5 in miniature: hundreds of small images in ONE launch sequence (one CTA per image, which also
    switches the frame-ring code to its thread-per-pixel mode)."""
    rows, cols, B = 96, 160, 320
    W = util.normal_w(rows, cols)
    wm = _mk(wmb, rows, cols, W)
    rng = np.random.default_rng(3)
    base = [util.natural_image(rows, cols, seed=600 + k) for k in range(8)]
    imgs = np.stack([np.clip(base[b % 8] + rng.uniform(-1, 1, (rows, cols)).astype(np.float32) * (b // 8), 0, 255) for b in range(B)]).astype(np.float32)
    L = wmb.lib()
    din = L.wm_dev_alloc(wm._h, imgs.nbytes)
    dout = L.wm_dev_alloc(wm._h, imgs.nbytes)
    L.wm_dev_upload(wm._h, din, imgs.ctypes.data, imgs.nbytes)
    n = rows * cols
    di = wmb.image_desc(din, rows, cols, wmb.ROW_MAJOR, wmb.F32)
    do = wmb.image_desc(dout, rows, cols, wmb.ROW_MAJOR, wmb.F32)
    a = np.zeros(B, np.float32)
    c = np.zeros(B, np.float32)
    st = np.zeros(B, np.int32)
    wm.embed_batch(0, di, di, do, n, n, n, B, wmb.ME, a, st)
    wm.detect_batch(0, do, n, B, wmb.ME, c, st)
    wm.sync(0)
    outs = np.zeros_like(imgs)
    L.wm_dev_download(wm._h, outs.ctypes.data, dout, outs.nbytes)
    for b in (0, 7, 100, B - 1):
        o = oracle.embed(imgs[b], W, 40.0, wmb.ME)
        od = oracle.detect(o["out"], W, wmb.ME)
        report("many_small b=%d a rel=%.3g dpix=%.3g corr rel=%.3g" % (
            b, abs(a[b] - o["a"]) / o["a"], np.abs(outs[b] - o["out"]).max(), abs(c[b] - od["corr"]) / abs(od["corr"])))
        assert abs(a[b] - o["a"]) / o["a"] <= 1e-3
        assert np.abs(outs[b] - o["out"]).max() <= 1e-4 * 255
        assert abs(c[b] - od["corr"]) / abs(od["corr"]) <= 1e-3
    # the same images one by one (many CTAs per image, few-pixels-per-block ring mode): the Rx partial sums are grouped differently, so
    # non-integer pixels may differ in the last bits of the f32 partials (the reference's own af::sum order is unspecified) ...
    for b in (3, 200):
        d1 = wmb.DeviceArray.from_numpy(wm, imgs[b], wmb.ROW_MAJOR)
        o1, a1, _ = wm.makeWatermark(d1, d1, wmb.ME)
        assert abs(a1 - a[b]) / a[b] <= 1e-5 and np.abs(o1.numpy() - outs[b]).max() <= 1e-4 * 255
    # ... while integer-valued pixels are summed exactly in every grouping: bit-identical scalars and pixels
    ints = np.rint(imgs[[3, 200]]).astype(np.float32)
    di2 = L.wm_dev_alloc(wm._h, ints.nbytes)
    do2 = L.wm_dev_alloc(wm._h, ints.nbytes)
    L.wm_dev_upload(wm._h, di2, ints.ctypes.data, ints.nbytes)
    a2 = np.zeros(2, np.float32)
    wm.embed_batch(0, wmb.image_desc(di2, rows, cols, wmb.ROW_MAJOR, wmb.F32), wmb.image_desc(di2, rows, cols, wmb.ROW_MAJOR, wmb.F32),
                   wmb.image_desc(do2, rows, cols, wmb.ROW_MAJOR, wmb.F32), n, n, n, 2, wmb.ME, a2)
    wm.sync(0)
    o2 = np.zeros_like(ints)
    L.wm_dev_download(wm._h, o2.ctypes.data, do2, o2.nbytes)
    for k in range(2):
        d1 = wmb.DeviceArray.from_numpy(wm, ints[k], wmb.ROW_MAJOR)
        o1, a1, _ = wm.makeWatermark(d1, d1, wmb.ME)
        assert a1 == a2[k] and np.array_equal(o1.numpy(), o2[k])
    L.wm_dev_free(wm._h, di2)
    L.wm_dev_free(wm._h, do2)
    # u8 frames through the same batched path: Rx exact => pixels identical to the oracle
    y = np.rint(imgs).astype(np.uint8)
    dy = L.wm_dev_alloc(wm._h, y.nbytes)
    dyo = L.wm_dev_alloc(wm._h, y.nbytes)
    L.wm_dev_upload(wm._h, dy, y.ctypes.data, y.nbytes)
    ctx = wmb.VideoProcessingContext(wm, rows, cols, 1, linesize=cols, frames_on_device=True)
    a8 = np.zeros(B, np.float32)
    wmb.process_frames(ctx, wmb.VIDEO_EMBED, dy, dyo, 0, B, a8)
    yo = np.zeros_like(y)
    L.wm_dev_download(wm._h, yo.ctypes.data, dyo, yo.nbytes)
    for b in (0, 150, B - 1):
        st8, oo, oa = oracle.embed_frame_u8(y[b], W, 40.0, oracle.ME)
        assert np.abs(yo[b].astype(int) - oo.astype(int)).max() <= 1 and abs(a8[b] - oa) / oa <= 1e-3
    for ptr in (din, dout, dy, dyo):
        L.wm_dev_free(wm._h, ptr)
    wm.close()


def test_8k_image(wmb, oracle):
    """BASELINE config 4: one 7680x4320 image (133 MB f32) — the reduction and the fused passes at HBM scale."""
    rows, cols = 4320, 7680
    img = util.natural_image(rows, cols, seed=8)
    W = util.normal_w(rows, cols)
    wm = _mk(wmb, rows, cols, W)
    sf = wm.strength_factor
    d = wmb.DeviceArray.from_numpy(wm, img, wmb.COL_MAJOR)
    for mask in (wmb.ME, wmb.NVF):
        out, a, st = wm.makeWatermark(d, d, mask)
        got = out.numpy()
        mse = float(np.mean((got.astype(np.float64) - img) ** 2))
        cm, _ = wm.detectWatermark(out, mask)
        cc, _ = wm.detectWatermark(d, mask)
        report("8k mask=%d a=%.7g mse/sf^2=%.6f corr marked=%.6f clean=%.6f" % (mask, a, mse / sf ** 2, cm, cc))
        assert st == 0 and 0.9 <= mse / sf ** 2 <= 1.0 + 1e-4
        assert cm > 0.2 and abs(cc) < 0.05
    # Rx of the 8K image against the oracle (33 Mpx of fp16-rounded products: the sums must still agree to 1e-6)
    wm.detectWatermark(d, wmb.ME)
    Rx = wm.debug(wmb.DBG_RX)
    oRx, _ = oracle.rx(img, oracle.FAITHFUL)
    report("8k Rx rel=%.3g" % util.rel(Rx, oRx))
    assert util.rel(Rx, oRx) <= 1e-6
    wm.close()


def test_rgb2gray_and_image_flow(wmb, oracle):
    """The reference's testForImage flow on the device: RGB -> gray (main.cpp:154), embed into the RGB image
    (main.cpp:178,190), gray of the watermarked RGB (main.cpp:196-197), detect (main.cpp:208,219)."""
    rgb = util.load_512_rgb()
    W = util.load_w512()
    wm = _mk(wmb, 512, 512, util.W512_PATH)
    for layout in LAYOUTS:
        drgb = wmb.DeviceArray.from_numpy(wm, rgb, layout)
        dg = wm.rgb2gray(drgb)
        g = dg.numpy()
        og = oracle.rgb2gray(rgb)
        assert np.array_equal(g, og)
        for mask in (wmb.NVF, wmb.ME):
            out, a, st = wm.makeWatermark(dg, drgb, mask)
            gw = wm.rgb2gray(out)
            corr, st2 = wm.detectWatermark(gw, mask)
            o = oracle.embed(og, W, 40.0, mask, base=rgb)
            od = oracle.detect(oracle.rgb2gray(o["out"]), W, mask)
            report("image_flow layout=%d mask=%d a=%.6f/%.6f corr=%.6f/%.6f" % (layout, mask, a, o["a"], corr, od["corr"]))
            assert st == 0 and st2 == 0
            assert abs(a - o["a"]) / o["a"] <= 1e-3 and abs(corr - od["corr"]) / abs(od["corr"]) <= 1e-3
    wm.close()


def test_clones_on_threads(wmb, oracle):
    """One object per concurrent caller (Watermark.cpp:30-34: copies own their workspace): clones used from several host
    threads at once give the same answers as a single caller."""
    import threading
    rows, cols = 256, 384
    W = util.normal_w(rows, cols)
    wm = _mk(wmb, rows, cols, W)
    imgs = [util.natural_image(rows, cols, seed=700 + k) for k in range(4)]
    ref = []
    for im in imgs:
        d = wmb.DeviceArray.from_numpy(wm, im, wmb.COL_MAJOR)
        out, a, _ = wm.makeWatermark(d, d, wmb.ME)
        c, _ = wm.detectWatermark(out, wmb.ME)
        ref.append((a, c))
    got = [None] * 4
    errs = []

    def work(k):
        try:
            w2 = wm.clone()
            d = wmb.DeviceArray.from_numpy(w2, imgs[k], wmb.COL_MAJOR)
            for _ in range(20):
                out, a, _ = w2.makeWatermark(d, d, wmb.ME)
                c, _ = w2.detectWatermark(out, wmb.ME)
            got[k] = (a, c)
            w2.close()
        except Exception as e:  # pragma: no cover
            errs.append(e)

    th = [threading.Thread(target=work, args=(k,)) for k in range(4)]
    for t in th:
        t.start()
    for t in th:
        t.join()
    assert not errs, errs
    assert got == ref
    wm.close()


@pytest.mark.parametrize("rows,cols", [(3, 3), (4, 5), (5, 7), (8, 300), (300, 8), (33, 129), (2, 64)])
def test_tiny_and_degenerate_shapes(wmb, oracle, rows, cols):
    """Far below the reference's 64-pixel minimum (main.cpp:161): the whole image is frame ring, tiles are mostly
    overhang.  Everything must still equal the oracle (or be rejected: fewer than 3 lines cannot hold a 3x3 core)."""
    W = util.normal_w(rows, cols)
    if rows < 3 or cols < 3:
        with pytest.raises(wmb.WatermarkError) as e:
            wmb.Watermark(rows, cols, W, 3, 40.0)
        assert e.value.code == -4
        return
    rng = np.random.default_rng(rows * 100 + cols)
    img = rng.integers(0, 256, (rows, cols)).astype(np.float32)  # white noise: well conditioned even when tiny
    wm = _mk(wmb, rows, cols, W)
    for layout in LAYOUTS:
        d = wmb.DeviceArray.from_numpy(wm, img, layout)
        corr, st = wm.detectWatermark(d, wmb.ME)
        Rx, rx = wm.debug(wmb.DBG_RX), wm.debug(wmb.DBG_RXVEC)
        oRx, orx = oracle.rx(img, oracle.FAITHFUL)
        assert np.array_equal(Rx, oRx) and np.array_equal(rx, orx)
        od = oracle.detect(img, W, wmb.ME)
        assert st == od["status"]
        if st == 0:
            assert abs(corr - od["corr"]) <= 1e-3 * max(abs(od["corr"]), 1e-2)
        for mask in (wmb.ME, wmb.NVF):
            out, a, st = wm.makeWatermark(d, d, mask)
            o = oracle.embed(img, W, 40.0, mask)
            assert st == o["status"]
            if st == 0:
                report("tiny %dx%d layout=%d mask=%d a rel=%.3g dpix=%.3g" % (rows, cols, layout, mask, abs(a - o["a"]) / o["a"], np.abs(out.numpy() - o["out"]).max()))
                assert abs(a - o["a"]) / o["a"] <= 1e-3
                assert np.abs(out.numpy() - o["out"]).max() <= 1e-4 * 255
    wm.close()


def test_sample_app_cli(wmb, oracle, tmp_path):
    """tools/sample_app.py: the reference's settings.ini-driven image flow (main.cpp:62-252) over the library — same keys,
    same printed results; correlations and strengths must match the oracle's run of the same flow."""
    import re
    import subprocess
    import sys
    ini = tmp_path / "settings.ini"
    ini.write_text("[paths]\nimage = %s\nwatermark = %s\n\n[options]\nsave_watermarked_files_to_disk = false\n"
                   "execution_time_in_fps = true\n\n[parameters]\np = 3\npsnr = 40.0\n; comment\nloops_for_test = 2\n"
                   % (util.PNG512_PATH, util.W512_PATH))
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, os.path.join(root, "tools", "sample_app.py"), str(ini)], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    out = r.stdout
    assert "Each test will be executed 2 times" in out and "Calculation of NVF mask with 512 rows and 512 columns" in out
    a = [float(x) for x in re.findall(r"Watermark strength \(parameter a\): ([0-9.eE+-]+)", out)]
    c_nvf = float(re.search(r"Correlation \[NVF\]: ([0-9.eE+-]+)", out).group(1))
    c_me = float(re.search(r"Correlation \[ME\]: ([0-9.eE+-]+)", out).group(1))
    rgb = util.load_512_rgb()
    W = util.load_w512()
    og = oracle.rgb2gray(rgb)
    for k, (mask, corr) in enumerate(((wmb.NVF, c_nvf), (wmb.ME, c_me))):
        o = oracle.embed(og, W, 40.0, mask, base=rgb)
        od = oracle.detect(oracle.rgb2gray(o["out"]), W, mask)
        report("sample_app mask=%d a=%.6f/%.6f corr=%.6f/%.6f" % (mask, a[k], o["a"], corr, od["corr"]))
        assert abs(a[k] - o["a"]) / o["a"] <= 1e-3 and abs(corr - od["corr"]) / abs(od["corr"]) <= 1e-3
    # the reference's argument checks (main.cpp:89-97)
    ini.write_text("[paths]\nimage = x.png\nwatermark = y\n[parameters]\np = 5\npsnr = 40\n")
    r = subprocess.run([sys.executable, os.path.join(root, "tools", "sample_app.py"), str(ini)], capture_output=True, text=True, timeout=120)
    assert r.returncode == 1 and "only p=3 is allowed" in r.stdout


@pytest.mark.parametrize("B,rows,cols", [(64, 512, 512), (150, 256, 384), (500, 64, 64)])
def test_unbalanced_batches_are_partitioned(wmb, oracle, B, rows, cols):
    """A batch whose size does not divide the resident CTA count is launched as sub-batches (wm_api.cu: partition).  With the
    split cost forced to 0 the partition is taken even for these small images; every image must come out as with a single
    launch (same scalars to f32 rounding of a different f64 summation grouping, same pixels) and match the oracle."""
    W = util.normal_w(rows, cols)
    wm = _mk(wmb, rows, cols, W)
    base_imgs = [util.natural_image(rows, cols, seed=900 + b) for b in range(8)]
    imgs = np.stack([np.roll(base_imgs[b % 8], (b // 8, 2 * (b // 8)), (0, 1)) for b in range(B)])
    if B > 70:
        imgs[69] = 23.0  # singular system in the second sub-batch
    L = wmb.lib()
    din = L.wm_dev_alloc(wm._h, imgs.nbytes)
    dout = L.wm_dev_alloc(wm._h, imgs.nbytes)
    L.wm_dev_upload(wm._h, din, imgs.ctypes.data, imgs.nbytes)
    npx = rows * cols
    di = wmb.image_desc(din, rows, cols, wmb.ROW_MAJOR, wmb.F32)
    do = wmb.image_desc(dout, rows, cols, wmb.ROW_MAJOR, wmb.F32)
    res = {}
    for cost in (-1, 0):
        wm.set_option(wmb.OPT_SPLIT_COST, cost)
        l0 = wm.launch_count
        for mask in (wmb.ME, wmb.NVF):
            a = np.full(B, np.nan, np.float32)
            st = np.zeros(B, np.int32)
            wm.embed_batch(0, di, di, do, npx, npx, npx, B, mask, a, st)
            wm.sync(0)
            outs = np.zeros_like(imgs)
            L.wm_dev_download(wm._h, outs.ctypes.data, dout, outs.nbytes)
            corr = np.zeros(B, np.float32)
            st2 = np.zeros(B, np.int32)
            wm.detect_batch(0, do, npx, B, mask, corr, st2)
            wm.sync(0)
            res[(cost, mask)] = (a, st.copy(), outs, corr, st2.copy())
        res[(cost, "launches")] = wm.launch_count - l0
    for mask in (wmb.ME, wmb.NVF):
        a0, s0, o0, c0, t0 = res[(-1, mask)]
        a1, s1, o1, c1, t1 = res[(0, mask)]
        assert np.array_equal(s0, s1) and np.array_equal(t0, t1)
        ok = s0 == 0
        if B > 70 and mask == wmb.ME:
            assert s0[69] != 0 and np.array_equal(o1[69], imgs[69])
        ra = np.abs(a1[ok] - a0[ok]) / np.abs(a0[ok])
        rc = np.abs(c1[ok] - c0[ok]) / np.abs(c0[ok])
        report("partition B=%d %dx%d mask=%d max rel a=%.3g corr=%.3g dpix=%.3g" % (B, rows, cols, mask, ra.max(), rc.max(), np.abs(o1 - o0).max()))
        assert ra.max() <= 1e-6 and rc.max() <= 1e-5 and np.abs(o1 - o0).max() <= 1e-4 * 255
        for b in (0, B // 2, B - 1):
            o = oracle.embed(imgs[b], W, 40.0, mask)
            od = oracle.detect(o["out"], W, mask)
            assert abs(a1[b] - o["a"]) / o["a"] <= 1e-3 and np.abs(o1[b] - o["out"]).max() <= 1e-4 * 255
            assert abs(c1[b] - od["corr"]) / abs(od["corr"]) <= 1e-3
    L.wm_dev_free(wm._h, din)
    L.wm_dev_free(wm._h, dout)
    wm.close()


@pytest.mark.parametrize("p", [5, 7, 9])
@pytest.mark.parametrize("rows,cols", [(64, 64), (67, 131), (270, 480)])
def test_nvf_larger_windows(wmb, oracle, rows, cols, p):
    """p in {5, 7, 9} (Watermark.cpp:24; kernels/nvf.hpp is generic in p): the NVF mask uses the p x p window, the
    prediction-error parts stay 3 x 3 as in the reference.  Mask planes bit-identical to the oracle (itself pinned against
    the reference's kernel text compiled with -Dp), embed / detect within north_star's tolerances, ME results untouched."""
    img = util.natural_image(rows, cols, seed=40 + p)
    W = util.normal_w(rows, cols)
    o = oracle.opts(p=p)
    wm = wmb.Watermark(rows, cols, W, p, 40.0)
    wm3 = _mk(wmb, rows, cols, W)
    onv = oracle.nvf(img, o)
    for layout in LAYOUTS:
        d = wmb.DeviceArray.from_numpy(wm, img, layout)
        nv = wm.debug_plane(d, wmb.DBG_MASK_NVF)
        report("nvf_p%d %dx%d layout=%d plane exact=%s rel=%.3g" % (p, rows, cols, layout, np.array_equal(nv, onv), util.rel(nv, onv)))
        assert np.array_equal(nv, onv)
        out, a, st = wm.makeWatermark(d, d, wmb.NVF)
        oe = oracle.embed(img, W, 40.0, wmb.NVF, o=o)
        got = out.numpy()
        assert st == 0 and abs(a - oe["a"]) / oe["a"] <= 1e-3 and np.abs(got - oe["out"]).max() <= 1e-4 * 255
        corr, st2 = wm.detectWatermark(wmb.DeviceArray.from_numpy(wm, got, layout), wmb.NVF)
        od = oracle.detect(got, W, wmb.NVF, o=o)
        report("nvf_p%d %dx%d layout=%d a=%.6f/%.6f corr=%.6f/%.6f" % (p, rows, cols, layout, a, oe["a"], corr, od["corr"]))
        assert st2 == 0 and abs(corr - od["corr"]) / abs(od["corr"]) <= 1e-3
        # the ME mask does not depend on p
        out_me, a_me, _ = wm.makeWatermark(d, d, wmb.ME)
        d3 = wmb.DeviceArray.from_numpy(wm3, img, layout)
        out3, a3, _ = wm3.makeWatermark(d3, d3, wmb.ME)
        assert a_me == a3 and np.array_equal(out_me.numpy(), out3.numpy())
    # u8 frames and a small batch through the same path
    y = util.natural_image(rows, cols, seed=50 + p, integer=True)
    dy = wmb.DeviceArray.from_numpy(wm, y, wmb.ROW_MAJOR)
    nvy = wm.debug_plane(dy, wmb.DBG_MASK_NVF)
    assert np.array_equal(nvy, oracle.nvf(y.astype(np.float32), o))
    B = 5
    imgs = np.stack([util.natural_image(rows, cols, seed=60 + b) for b in range(B)])
    L = wmb.lib()
    din = L.wm_dev_alloc(wm._h, imgs.nbytes)
    dout = L.wm_dev_alloc(wm._h, imgs.nbytes)
    L.wm_dev_upload(wm._h, din, imgs.ctypes.data, imgs.nbytes)
    di = wmb.image_desc(din, rows, cols, wmb.ROW_MAJOR, wmb.F32)
    do = wmb.image_desc(dout, rows, cols, wmb.ROW_MAJOR, wmb.F32)
    ab = np.zeros(B, np.float32)
    cb = np.zeros(B, np.float32)
    wm.embed_batch(0, di, di, do, rows * cols, rows * cols, rows * cols, B, wmb.NVF, ab)
    wm.detect_batch(0, do, rows * cols, B, wmb.NVF, cb)
    wm.sync(0)
    for b in (0, B - 1):
        oe = oracle.embed(imgs[b], W, 40.0, wmb.NVF, o=o)
        od = oracle.detect(oe["out"], W, wmb.NVF, o=o)
        assert abs(ab[b] - oe["a"]) / oe["a"] <= 1e-3 and abs(cb[b] - od["corr"]) / abs(od["corr"]) <= 1e-3
    L.wm_dev_free(wm._h, din)
    L.wm_dev_free(wm._h, dout)
    wm.close()
    wm3.close()


@pytest.mark.parametrize("rows,cols", [(3, 3), (4, 5), (8, 300), (33, 129)])
def test_nvf_larger_windows_tiny_and_strided(wmb, oracle, rows, cols):
    """p = 9 on images smaller than the window (every read clamps) and on a strided buffer."""
    p = 9
    img = util.natural_image(rows, cols, seed=70 + rows)
    W = util.normal_w(rows, cols)
    o = oracle.opts(p=p)
    wm = wmb.Watermark(rows, cols, W, p, 40.0)
    onv = oracle.nvf(img, o)
    for layout in LAYOUTS:
        for ld in (0, (cols if layout == wmb.ROW_MAJOR else rows) + 5):
            d = wmb.DeviceArray.from_numpy(wm, img, layout, ld=ld)
            assert np.array_equal(wm.debug_plane(d, wmb.DBG_MASK_NVF), onv)
            out, a, st = wm.makeWatermark(d, d, wmb.NVF)
            oe = oracle.embed(img, W, 40.0, wmb.NVF, o=o)
            if oe["status"] == 0 and np.isfinite(oe["a"]):
                assert st == 0 and abs(a - oe["a"]) / abs(oe["a"]) <= 1e-3
                assert np.abs(out.numpy() - oe["out"]).max() <= 1e-4 * 255
    wm.close()
