"""GPU tests added in round 2: strided host views, the pipelined host batch API, batch-stride conventions, the multi-GPU video driver
(contexts on one device stand in for several GPUs), embed + verify mode, the f32 solve option and the spread between the oracle's
precision modes (the reference sums and solves in f32: Watermark.cpp:148-149,203)."""
import ctypes as C

import numpy as np
import pytest

import util
from test_gpu_parity import _mk, report

pytestmark = pytest.mark.gpu


# ---------------------------------------------------------------------------------------------------
# host-buffer API on strided views: only the images' own pixels may be read or written
# ---------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("layout", [0, 1])
def test_host_api_strided_views_keep_padding(wmb, oracle, layout):
    rows, cols, pad = 72, 100, 12
    img = util.natural_image(rows, cols, seed=41)
    rgb = np.stack([np.clip(img + d, 0, 255) for d in (-7.0, 0.0, 9.0)]).astype(np.float32)
    W = util.normal_w(rows, cols)
    wm = _mk(wmb, rows, cols, W)
    L_, P_ = (cols, rows) if layout == wmb.COL_MAJOR else (rows, cols)
    ld = P_ + pad
    SENT = np.float32(-12345.0)

    def to_mem(planes):  # (ch, rows, cols) -> strided memory (ch, L + 2 lines of plane padding, ld)
        ch = planes.shape[0]
        buf = np.full((ch, L_ + 2, ld), SENT, np.float32)
        for c in range(ch):
            buf[c, :L_, :P_] = planes[c].T if layout == wmb.COL_MAJOR else planes[c]
        return buf

    def from_mem(buf):
        v = buf[:, :L_, :P_]
        return np.stack([p.T if layout == wmb.COL_MAJOR else p for p in v])

    gin = to_mem(img[None])
    base = to_mem(rgb)
    out = np.full_like(base, SENT)
    ps = (L_ + 2) * ld
    di = wmb.image_desc(gin.ctypes.data, rows, cols, layout, wmb.F32, ld=ld, channels=1, plane_stride=ps)
    db = wmb.image_desc(base.ctypes.data, rows, cols, layout, wmb.F32, ld=ld, channels=3, plane_stride=ps)
    do = wmb.image_desc(out.ctypes.data, rows, cols, layout, wmb.F32, ld=ld, channels=3, plane_stride=ps)
    a = C.c_float(0)
    for mask in (wmb.ME, wmb.NVF):
        out[:] = SENT
        rc = wmb.lib().wm_embed_host(wm._h, C.byref(di), C.byref(db), C.byref(do), mask, C.byref(a))
        assert rc == 0
        o = oracle.embed(img, W, 40.0, mask, base=rgb)
        got = from_mem(out)
        assert abs(a.value - o["a"]) / o["a"] <= 1e-3
        assert np.abs(got - o["out"]).max() <= 1e-4 * 255
        # every byte outside the three planes' pixels is untouched
        chk = out.copy()
        chk[:, :L_, :P_] = SENT
        assert np.all(chk == SENT), "row / plane padding of the output view was overwritten"
        # detection from a strided one-channel host view of the first plane
        dz = wmb.image_desc(out.ctypes.data, rows, cols, layout, wmb.F32, ld=ld, channels=1, plane_stride=ps)
        corr = C.c_float(0)
        assert wmb.lib().wm_detect_host(wm._h, C.byref(dz), mask, C.byref(corr)) == 0
        od = oracle.detect(o["out"][0], W, mask)
        assert abs(corr.value - od["corr"]) / abs(od["corr"]) <= 1e-3
    assert np.all(gin[:, L_:, :] == SENT) and np.all(gin[:, :, P_:] == SENT)
    wm.close()


def test_host_batch_api_pipelined(wmb, oracle):
    rows, cols, B = 96, 128, 6
    W = util.normal_w(rows, cols)
    wm = _mk(wmb, rows, cols, W)
    imgs = np.stack([util.natural_image(rows, cols, seed=500 + b) for b in range(B)])
    outs = np.zeros_like(imgs)
    a = np.zeros(B, np.float32)
    corr = np.zeros(B, np.float32)
    st = np.zeros(B, np.int32)
    npx = rows * cols
    # two calls of 3 frames on two slots, then detection of the watermarked HOST frames on the same slots (stream order makes the
    # D2H copy of `outs` finish before the H2D copy that re-reads it)
    for k, sl in ((0, 1), (1, 2)):
        o = 3 * k
        hin = wmb.image_desc(imgs[o].ctypes.data, rows, cols, wmb.ROW_MAJOR, wmb.F32)
        hout = wmb.image_desc(outs[o].ctypes.data, rows, cols, wmb.ROW_MAJOR, wmb.F32)
        wm.embed_host_batch(sl, hin, hin, hout, npx, npx, npx, 3, wmb.ME, a[o:o + 3], st[o:o + 3])
        wm.detect_host_batch(sl, hout, npx, 3, wmb.ME, corr[o:o + 3])
    wm.sync(-1)
    for b in range(B):
        o = oracle.embed(imgs[b], W, 40.0, wmb.ME)
        od = oracle.detect(o["out"], W, wmb.ME)
        assert st[b] == 0
        assert abs(a[b] - o["a"]) / o["a"] <= 1e-3
        assert np.abs(outs[b] - o["out"]).max() <= 1e-4 * 255
        assert abs(corr[b] - od["corr"]) / abs(od["corr"]) <= 1e-3
    wm.close()


# ---------------------------------------------------------------------------------------------------
# batch strides: 0 = dense; overlapping images and out-over-in are rejected
# ---------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("tma", [1, 0])
def test_batch_stride_zero_is_dense_and_overlaps_are_rejected(wmb, oracle, tma):
    rows, cols, B = 64, 128, 4
    W = util.normal_w(rows, cols)
    wm = _mk(wmb, rows, cols, W)
    wm.set_option(wmb.OPT_USE_TMA, tma)
    imgs = np.stack([util.natural_image(rows, cols, seed=600 + b) for b in range(B)])
    L = wmb.lib()
    din = L.wm_dev_alloc(wm._h, imgs.nbytes)
    dout = L.wm_dev_alloc(wm._h, imgs.nbytes)
    L.wm_dev_upload(wm._h, din, imgs.ctypes.data, imgs.nbytes)
    di = wmb.image_desc(din, rows, cols, wmb.ROW_MAJOR, wmb.F32)
    do = wmb.image_desc(dout, rows, cols, wmb.ROW_MAJOR, wmb.F32)
    npx = rows * cols
    res = {}
    for stride in (npx, 0):
        a = np.zeros(B, np.float32)
        wm.embed_batch(0, di, di, do, stride, stride, stride, B, wmb.ME, a)
        wm.sync(0)
        o = np.zeros_like(imgs)
        L.wm_dev_download(wm._h, o.ctypes.data, dout, o.nbytes)
        c = np.zeros(B, np.float32)
        wm.detect_batch(0, do, stride, B, wmb.ME, c)
        wm.sync(0)
        res[stride] = (a, o, c)
    assert np.array_equal(res[0][0], res[npx][0]) and np.array_equal(res[0][1], res[npx][1]) and np.array_equal(res[0][2], res[npx][2])
    assert len(set(res[0][0].tolist())) == B  # four different images gave four different strengths
    a = np.zeros(B, np.float32)
    with pytest.raises(wmb.WatermarkError) as e:   # images of the batch would overlap
        wm.embed_batch(0, di, di, do, npx // 2, npx // 2, npx, B, wmb.ME, a)
    assert e.value.code == -5
    with pytest.raises(wmb.WatermarkError):        # out inside the input range
        do2 = wmb.image_desc(din + 4 * npx, rows, cols, wmb.ROW_MAJOR, wmb.F32)
        wm.embed_batch(0, di, di, do2, npx, npx, npx, 2, wmb.ME, a)
    with pytest.raises(wmb.WatermarkError):
        wm.detect_batch(0, di, 16, B, wmb.ME, a)
    L.wm_dev_free(wm._h, din)
    L.wm_dev_free(wm._h, dout)
    wm.close()


# ---------------------------------------------------------------------------------------------------
# multi-GPU video driver: contiguous chunks of the global index, one host thread per context
# ---------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("on_device", [False, True])
@pytest.mark.parametrize("ngpus,interval", [(2, 1), (3, 2), (4, 3)])
def test_process_frames_multi_equals_single(wmb, oracle, ngpus, interval, on_device):
    rows, cols, ls, n, first = 96, 160, 176, 11, 5
    W = util.normal_w(rows, cols)
    wm = _mk(wmb, rows, cols, W)
    clones = [wm.clone() for _ in range(ngpus)]  # one context per "device" (all on device 0 here: the chunking and threading are the same)
    frames = np.zeros((n, rows, ls), np.uint8)
    for i in range(n):
        frames[i, :, :cols] = util.natural_image(rows, cols, seed=700 + i, integer=True)
    L = wmb.lib()
    fbytes, obytes = rows * ls, rows * cols
    if on_device:
        dfr = L.wm_dev_alloc(wm._h, frames.nbytes)
        dout = L.wm_dev_alloc(wm._h, n * obytes)
        dout2 = L.wm_dev_alloc(wm._h, n * obytes)
        L.wm_dev_upload(wm._h, dfr, frames.ctypes.data, frames.nbytes)
        src, dst, dst2 = dfr, dout, dout2
    else:
        out_h, out_h2 = np.zeros((n, rows, cols), np.uint8), np.zeros((n, rows, cols), np.uint8)
        src, dst, dst2 = frames.ctypes.data, out_h.ctypes.data, out_h2.ctypes.data

    def download(ptr, host):
        if on_device:
            o = np.zeros((n, rows, cols), np.uint8)
            L.wm_dev_download(wm._h, o.ctypes.data, ptr, o.nbytes)
            return o
        return host.copy()

    one = wmb.VideoProcessingContext(wm, rows, cols, interval, linesize=ls, frames_on_device=on_device)
    a1 = np.zeros(n, np.float32)
    wmb.process_frames(one, wmb.VIDEO_EMBED, src, dst, first, n, a1)
    ref_out = download(dst, None if on_device else out_h)
    ctxs = [wmb.VideoProcessingContext(c, rows, cols, interval, linesize=ls, frames_on_device=on_device) for c in clones]
    firsts = [wmb.shard_frames(n, g, ngpus)[0] for g in range(ngpus)]
    a2 = np.full(n, 7.0, np.float32)
    got = wmb.process_frames_multi(ctxs, wmb.VIDEO_EMBED, [src + f * fbytes for f in firsts], [dst2 + f * obytes for f in firsts], first, n, a2)
    assert got == n
    out2 = download(dst2, None if on_device else out_h2)
    assert np.array_equal(ref_out, out2)
    assert np.array_equal(a1, a2, equal_nan=True)
    gated = (first + np.arange(n)) % interval == 0
    assert np.all(np.isnan(a2[~gated])) and not np.any(np.isnan(a2[gated]))
    # detection over the chunks, same comparison
    one_d = wmb.VideoProcessingContext(wm, rows, cols, interval, linesize=cols, frames_on_device=on_device)
    c1, c2 = np.zeros(n, np.float32), np.zeros(n, np.float32)
    wmb.process_frames(one_d, wmb.VIDEO_DETECT, dst, None, first, n, c1)
    ctxs_d = [wmb.VideoProcessingContext(c, rows, cols, interval, linesize=cols, frames_on_device=on_device) for c in clones]
    wmb.process_frames_multi(ctxs_d, wmb.VIDEO_DETECT, [dst2 + f * obytes for f in firsts], None, first, n, c2)
    assert np.array_equal(c1, c2, equal_nan=True)
    i = int(np.flatnonzero(gated)[-1])
    st, oo, oa = oracle.embed_frame_u8(frames[i], W, 40.0, oracle.ME, width=cols)
    assert abs(a2[i] - oa) / oa <= 1e-3 and np.abs(out2[i].astype(int) - oo.astype(int)).max() <= 1
    if on_device:
        for p in (dfr, dout, dout2):
            L.wm_dev_free(wm._h, p)
    for c in clones:
        c.close()
    wm.close()


def test_shard_frames_matches_python_split(wmb):
    for n in (1, 7, 64, 513, 1036):
        for world in (1, 2, 3, 4, 8):
            nxt = 0
            for r in range(world):
                first, cnt = wmb.shard_frames(n, r, world)
                base, rem = divmod(n, world)
                assert first == r * base + min(r, rem) and cnt == base + (1 if r < rem else 0)
                assert first == nxt
                nxt += cnt
            assert nxt == n


@pytest.mark.parametrize("on_device", [False, True])
def test_video_embed_verify_mode(wmb, oracle, on_device):
    """EMBED_VERIFY = EMBED followed by DETECT on the written frames, without a second upload."""
    rows, cols, ls, n, interval, first = 120, 200, 224, 9, 2, 3
    W = util.normal_w(rows, cols)
    wm = _mk(wmb, rows, cols, W)
    frames = np.zeros((n, rows, ls), np.uint8)
    for i in range(n):
        frames[i, :, :cols] = util.natural_image(rows, cols, seed=800 + i, integer=True)
    frames[4, :, :cols] = 93  # global index 7: gated off (interval 2)
    frames[5, :, :cols] = 93  # global index 8: gated on, constant -> unsolvable: a stays NaN, frame copied through
    L = wmb.lib()
    out_a, out_b = np.zeros((n, rows, cols), np.uint8), np.zeros((n, rows, cols), np.uint8)
    if on_device:
        dfr = L.wm_dev_alloc(wm._h, frames.nbytes)
        da, db = L.wm_dev_alloc(wm._h, out_a.nbytes), L.wm_dev_alloc(wm._h, out_b.nbytes)
        L.wm_dev_upload(wm._h, dfr, frames.ctypes.data, frames.nbytes)
        src, pa, pb = dfr, da, db
    else:
        src, pa, pb = frames.ctypes.data, out_a.ctypes.data, out_b.ctypes.data
    ctx = wmb.VideoProcessingContext(wm, rows, cols, interval, linesize=ls, frames_on_device=on_device)
    ctx_d = wmb.VideoProcessingContext(wm, rows, cols, interval, linesize=cols, frames_on_device=on_device)
    a1, c1 = np.zeros(n, np.float32), np.zeros(n, np.float32)
    wmb.process_frames(ctx, wmb.VIDEO_EMBED, src, pa, first, n, a1)
    wmb.process_frames(ctx_d, wmb.VIDEO_DETECT, pa, None, first, n, c1)
    both = np.zeros(2 * n, np.float32)
    wmb.process_frames(ctx, wmb.VIDEO_EMBED_VERIFY, src, pb, first, n, both)
    if on_device:
        L.wm_dev_download(wm._h, out_a.ctypes.data, da, out_a.nbytes)
        L.wm_dev_download(wm._h, out_b.ctypes.data, db, out_b.nbytes)
    assert np.array_equal(out_a, out_b)
    assert np.array_equal(a1, both[:n], equal_nan=True) and np.array_equal(c1, both[n:], equal_nan=True)
    assert np.isnan(a1[5]) and np.array_equal(out_a[5], frames[5, :, :cols]) and c1[5] == 0.0   # unsolvable frame (Watermark.cpp:164-165,246-247)
    assert np.isnan(a1[4]) and np.isnan(c1[4])                                                    # gated off
    assert not np.isnan(a1[1]) and not np.isnan(c1[1])
    if on_device:
        for p in (dfr, da, db):
            L.wm_dev_free(wm._h, p)
    wm.close()


# ---------------------------------------------------------------------------------------------------
# precision modes: the reference reduces and solves in f32 (ArrayFire); ours in f64.  Measure and bound the spread.
# ---------------------------------------------------------------------------------------------------
def test_f32_solve_option_matches_f32_lu(wmb, oracle):
    img = util.load_512_gray(oracle)
    W = util.load_w512()
    wm = _mk(wmb, 512, 512, W)
    d = wmb.DeviceArray.from_numpy(wm, img, wmb.COL_MAJOR)
    wm.detectWatermark(d, wmb.ME)
    c64 = wm.debug(wmb.DBG_COEFFS).copy()
    Rx, rx = wm.debug(wmb.DBG_RX), wm.debug(wmb.DBG_RXVEC)
    wm.set_option(wmb.OPT_F32_SOLVE, 1)
    wm.detectWatermark(d, wmb.ME)
    c32 = wm.debug(wmb.DBG_COEFFS).copy()
    st, o32 = oracle.solve8(Rx, rx, f32=True)
    st, o64 = oracle.solve8(Rx, rx, f32=False)
    report("f32-solve option: |c32 - oracle f32 LU| = %.3g, |c32 - c64| = %.3g (rel to max|c|)" % (
        util.rel(c32, o32), util.rel(c32, c64)))
    assert np.array_equal(c64, o64.astype(np.float32))
    assert util.rel(c32, o32) <= 2e-6          # same algorithm in f32; the host LU may contract mul-sub into fma
    assert 0 < util.rel(c32, c64) <= 5e-3      # the f32 LU of this cond ~1e4 system moves the coefficients visibly
    wm.close()


@pytest.mark.parametrize("name", ["512", "1080p"])
def test_gpu_lies_inside_the_spread_of_the_precision_modes(wmb, oracle, name):
    """run the oracle as the reference would (f32 sums + f32 LU, STRICT_F32), canonically (f64) and without the fp16
    rounding (EXACT); print how far those differ from each other and how far the GPU is from the canonical mode.  The GPU must be
    much closer to the canonical mode than the modes are to each other — i.e. it sits inside the reference's own noise."""
    if name == "512":
        img, W = util.load_512_gray(oracle), util.load_w512()
    else:
        img, W = util.natural_image(1080, 1920, seed=5), util.normal_w(1080, 1920, seed=28390211)
    rows, cols = img.shape
    wm = _mk(wmb, rows, cols, W)
    d = wmb.DeviceArray.from_numpy(wm, img, wmb.COL_MAJOR)
    for mask in (wmb.ME, wmb.NVF):
        out, a, st = wm.makeWatermark(d, d, mask)
        coef = wm.debug(wmb.DBG_COEFFS).copy()
        got = out.numpy()
        corr, st2 = wm.detectWatermark(out, mask)
        modes = {"canonical": oracle.FAITHFUL, "strict_f32": oracle.STRICT_F32, "exact": oracle.EXACT}
        ref = {}
        for k, o in modes.items():
            e = oracle.embed(img, W, 40.0, mask, o=o)
            dd = oracle.detect(e["out"], W, mask, o=o)
            ref[k] = dict(a=e["a"], corr=dd["corr"], out=e["out"], coef=e["coef"], mask=e["mask"])
        can = ref["canonical"]
        gpu_dev = dict(a=abs(a - can["a"]) / abs(can["a"]), corr=abs(corr - can["corr"]) / abs(can["corr"]),
                       out=float(np.abs(got - can["out"]).max()) / 255.0,
                       coef=util.rel(coef, can["coef"]) if mask == wmb.ME else 0.0)
        for k in ("strict_f32", "exact"):
            r = ref[k]
            spread = dict(a=abs(r["a"] - can["a"]) / abs(can["a"]), corr=abs(r["corr"] - can["corr"]) / abs(can["corr"]),
                          out=float(np.abs(r["out"] - can["out"]).max()) / 255.0,
                          coef=util.rel(r["coef"], can["coef"]) if mask == wmb.ME else 0.0, mask=util.rel(r["mask"], can["mask"]))
            report("spread %s mask=%d %-10s vs canonical: a %.3g corr %.3g pixels/255 %.3g coef %.3g mask %.3g | GPU vs canonical: a %.3g corr %.3g pixels/255 %.3g coef %.3g" % (
                name, mask, k, spread["a"], spread["corr"], spread["out"], spread["coef"], spread["mask"],
                gpu_dev["a"], gpu_dev["corr"], gpu_dev["out"], gpu_dev["coef"]))
            for q in ("a", "corr", "out", "coef"):
                assert gpu_dev[q] <= max(spread[q], 2e-6), (q, gpu_dev[q], spread[q])
        assert gpu_dev["a"] <= 1e-3 and gpu_dev["corr"] <= 1e-3 and gpu_dev["out"] <= 1e-4
    wm.close()


# ---------------------------------------------------------------------------------------------------
# the detector's three sums one by one, and NVF detection where every pixel touches the nested clamp
# ---------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("layout", [0, 1])
@pytest.mark.parametrize("rows,cols", [(3, 3), (4, 5), (5, 7), (8, 300), (300, 8), (33, 129), (64, 64), (130, 260)])
def test_detector_sums_individually(wmb, oracle, rows, cols, layout):
    img = util.natural_image(rows, cols, seed=900 + rows + cols)
    W = util.normal_w(rows, cols)
    wm = _mk(wmb, rows, cols, W)
    d = wmb.DeviceArray.from_numpy(wm, img, layout)
    for mask in (wmb.ME, wmb.NVF):
        od = oracle.detect(img, W, mask)
        if od["status"] != 0:
            continue
        wm.debug_set_coeffs(od["coef"])  # same coefficients on both sides: the sums are compared, not the solve
        corr, st = wm.detectWatermark(d, mask)
        s = wm.debug(wmb.DBG_SCALARS)
        ez, eu, u = od["ez"].astype(np.float64), od["eu"].astype(np.float64), od["u"].astype(np.float64)
        # the kernel drops the 1 / max|e| scale of the ME mask (it cancels in the correlation): rescale its sums to compare
        k = float(np.abs(od["ez"]).max()) if mask == wmb.ME else 1.0
        want = dict(dot=float((eu * ez).sum()), nz=float((ez * ez).sum()), nu=float((eu * eu).sum()))
        got = dict(dot=s[4] / k, nz=s[5], nu=s[6] / (k * k))
        for q in want:
            rel = abs(got[q] - want[q]) / max(abs(want[q]), 1e-30)
            report("detect sums %dx%d layout=%d mask=%d %s rel=%.3g" % (rows, cols, layout, mask, q, rel))
            assert rel <= 2e-5, (q, got[q], want[q])
        assert abs(corr - od["corr"]) / max(abs(od["corr"]), 1e-30) <= 1e-4
        wm.debug_set_coeffs(None)
    wm.close()


# ---------------------------------------------------------------------------------------------------
# the detector's never-materialised planes (u, e_u) and the ME mask, bit for bit with injected coefficients (MASK, U, EU)
# ---------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("layout", [0, 1])
@pytest.mark.parametrize("rows,cols", [(3, 3), (5, 7), (33, 129), (64, 64), (67, 131), (130, 260), (512, 512)])
def test_detector_planes_bit_exact(wmb, oracle, rows, cols, layout):
    img = util.natural_image(rows, cols, seed=1000 + rows * 3 + cols) if rows != 512 else util.load_512_gray(oracle)
    W = util.normal_w(rows, cols) if rows != 512 else util.load_w512()
    wm = _mk(wmb, rows, cols, W)
    d = wmb.DeviceArray.from_numpy(wm, img, layout)
    for mask in (wmb.NVF, wmb.ME):
        od = oracle.detect(img, W, mask)
        if od["status"] != 0:
            continue
        wm.debug_set_coeffs(od["coef"])
        u, eu, corr = wm.debug_detect_planes(d, mask)
        if mask == wmb.NVF:
            want_u = od["u"]
        else:
            want_u = (np.abs(od["ez"]) * W).astype(np.float32)  # the kernel keeps |e_z|.W: the 1 / max|e| scale cancels in the correlation
        want_eu = (want_u - oracle.scaled_neighbors(want_u, od["coef"])).astype(np.float32)
        report("detector planes %dx%d layout=%d mask=%d: u maxdiff %.3g, e_u maxdiff %.3g" % (
            rows, cols, layout, mask, np.abs(u - want_u).max(), np.abs(eu - want_eu).max()))
        assert np.array_equal(u, want_u)
        assert np.array_equal(eu, want_eu)
        if mask == wmb.NVF:
            assert np.array_equal(eu, od["eu"])
        assert abs(corr - od["corr"]) / max(abs(od["corr"]), 1e-30) <= 1e-4
        if mask == wmb.ME:
            pm = oracle.pred_error_mask(img)
            m = wm.debug_plane(d, wmb.DBG_MASK_ME)
            assert np.array_equal(m, pm["mask"])
        wm.debug_set_coeffs(None)
    wm.close()


# ---------------------------------------------------------------------------------------------------
# the apply kernel's two output paths — TMA stores (WM_OPT_TMA_STORE, default) and per-thread vector stores — must give the same bits
# ---------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("dtype", ["f32", "u8"])
@pytest.mark.parametrize("rows,cols,B", [(64, 128, 1), (96, 160, 5), (200, 384, 3), (270, 480, 2), (1080, 1920, 2)])
def test_tma_store_variant_same_bits(wmb, oracle, rows, cols, B, dtype):
    W = util.normal_w(rows, cols)
    wm = _mk(wmb, rows, cols, W)
    imgs = np.stack([util.natural_image(rows, cols, seed=1100 + b, integer=(dtype == "u8")) for b in range(B)])
    if dtype == "f32":
        imgs = imgs.astype(np.float32)
    dt = wmb.U8 if dtype == "u8" else wmb.F32
    L = wmb.lib()
    din = L.wm_dev_alloc(wm._h, imgs.nbytes)
    dout = L.wm_dev_alloc(wm._h, imgs.nbytes)
    L.wm_dev_upload(wm._h, din, imgs.ctypes.data, imgs.nbytes)
    di = wmb.image_desc(din, rows, cols, wmb.ROW_MAJOR, dt)
    do = wmb.image_desc(dout, rows, cols, wmb.ROW_MAJOR, dt)
    n = rows * cols
    res = []
    for ts in (0, 1):
        wm.set_option(wmb.OPT_TMA_STORE, ts)
        for mask in (wmb.ME, wmb.NVF):
            sent = np.full_like(imgs, 7)
            L.wm_dev_upload(wm._h, dout, sent.ctypes.data, sent.nbytes)
            a = np.zeros(B, np.float32)
            wm.embed_batch(0, di, di, do, n, n, n, B, mask, a)
            wm.sync(0)
            o = np.zeros_like(imgs)
            L.wm_dev_download(wm._h, o.ctypes.data, dout, o.nbytes)
            res.append((a, o))
    for k in range(2):
        assert np.array_equal(res[k][0], res[2 + k][0]) and np.array_equal(res[k][1], res[2 + k][1])
    assert not np.array_equal(res[0][1], imgs)
    L.wm_dev_free(wm._h, din)
    L.wm_dev_free(wm._h, dout)
    wm.close()


def test_host_embed_verify_batch(wmb, oracle):
    rows, cols, B = 96, 128, 4
    W = util.normal_w(rows, cols)
    wm = _mk(wmb, rows, cols, W)
    imgs = np.stack([util.natural_image(rows, cols, seed=1300 + b) for b in range(B)])
    npx = rows * cols
    for mask in (wmb.NVF, wmb.ME):
        outs = np.zeros_like(imgs)
        a, c, a2, c2 = (np.zeros(B, np.float32) for _ in range(4))
        hin = wmb.image_desc(imgs.ctypes.data, rows, cols, wmb.ROW_MAJOR, wmb.F32)
        hout = wmb.image_desc(outs.ctypes.data, rows, cols, wmb.ROW_MAJOR, wmb.F32)
        wm.embed_verify_host_batch(3, hin, hin, hout, npx, npx, npx, B, mask, a, c)
        wm.sync(3)
        outs2 = np.zeros_like(imgs)
        hout2 = wmb.image_desc(outs2.ctypes.data, rows, cols, wmb.ROW_MAJOR, wmb.F32)
        wm.embed_host_batch(4, hin, hin, hout2, npx, npx, npx, B, mask, a2)
        wm.detect_host_batch(4, hout2, npx, B, mask, c2)
        wm.sync(4)
        assert np.array_equal(outs, outs2) and np.array_equal(a, a2) and np.array_equal(c, c2)
        o = oracle.embed(imgs[1], W, 40.0, mask)
        assert abs(a[1] - o["a"]) / o["a"] <= 1e-3 and abs(c[1] - oracle.detect(o["out"], W, mask)["corr"]) / abs(c[1]) <= 1e-3
    wm.close()


# ---------------------------------------------------------------------------------------------------
# single-image fused kernels (one cooperative launch per synchronous op) == the multi-kernel path
# ---------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("rows,cols", [(1080, 1920), (512, 512), (700, 1000)])
@pytest.mark.parametrize("layout", [0, 1])
def test_fused_single_image_ops_equal_multi_kernel_path(wmb, oracle, rows, cols, layout):
    """wm_embed / wm_detect on one resident f32 image, called repeatedly (the reference's loops_for_test protocol, main.cpp:167-223): from the
    second call on the op is captured / replayed and, with WM_OPT_FUSED_SINGLE = 1, wm_detect runs as ONE cooperative kernel (k_detect1).  Integer-valued pixels: every sum is exact,
    so the fused op must give the same bits as the default multi-kernel path; the launch counter proves which path ran."""
    img = util.natural_image(rows, cols, seed=11, integer=True).astype(np.float32)
    W = util.normal_w(rows, cols)
    res = {}
    for fused in (1, 0):
        wm = wmb.Watermark(rows, cols, W, 3, 40.0)
        wm.set_option(wmb.OPT_FUSED_SINGLE, fused)
        d = wmb.DeviceArray.from_numpy(wm, img, layout)
        out = wmb.DeviceArray(wm, rows, cols, layout, wmb.F32)
        r = {}
        for mask in (wmb.ME, wmb.NVF):
            for _ in range(4):  # 1st: plain, 2nd: capture, 3rd/4th: replay
                n0 = wm.launch_count
                _, a, st = wm.makeWatermark(d, d, mask, out=out)
                ne = wm.launch_count - n0
            assert st == 0
            z = out.numpy().copy()
            for _ in range(4):
                n0 = wm.launch_count
                corr, st = wm.detectWatermark(out, mask)
                nd = wm.launch_count - n0
            assert st == 0
            r[mask] = (a, corr, z, ne, nd, wm.debug(wmb.DBG_COEFFS).copy())
        res[fused] = r
        wm.close()
    for mask in (wmb.ME, wmb.NVF):
        a1, c1, z1, ne1, nd1, k1 = res[1][mask]
        a0, c0, z0, ne0, nd0, k0 = res[0][mask]
        assert nd0 == 2 and ne0 == (3 if mask == wmb.ME else 2)
        assert nd1 == 1, "the fused detector did not run"
        assert np.array_equal(k1, k0) and c1 == c0, (mask, c1, c0)
        assert a1 == a0 and np.array_equal(z1, z0)
        o = oracle.detect(z1, W, mask)
        assert abs(c1 - o["corr"]) <= 1e-3 * abs(o["corr"])
        report("fused single-image ops %dx%d layout=%d mask=%d: corr %.7f (multi-kernel %.7f, oracle %.7f), launches embed %d detect %d" % (
            rows, cols, layout, mask, c1, c0, o["corr"], ne1, nd1))


# ---------------------------------------------------------------------------------------------------
# host video frames: uploaded with their row padding (one linear copy, read in place) == repacked on the way up (2-D copy)
# ---------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("interval", [1, 2])
@pytest.mark.parametrize("ls_extra,frame_extra", [(64, 0), (64, 4096), (16, 0), (0, 0), (48 + 640, 0)])
def test_host_frames_padded_upload_equals_repack(wmb, oracle, interval, ls_extra, frame_extra):
    """main.cpp:348-353 repacks every decoded frame row by row to drop ffmpeg's padding.  wm_process_frames either drops it with a 2-D copy
    (WM_OPT_PADDED_UPLOAD = 0) or — small paddings, 16-byte aligned rows — uploads the frame with it and reads it in place.  Same frames, same
    scalars, bit for bit; frame strides with trailing chroma planes (frame_extra) and a padding too large to keep (last case) included."""
    rows, cols, n, first = 136, 640, 7, 5
    ls = cols + ls_extra
    fstride = rows * ls + frame_extra
    W = util.normal_w(rows, cols)
    buf = np.full(n * fstride, 201, np.uint8)  # padding and chroma bytes hold junk that must never be read as pixels
    for i in range(n):
        fr = buf[i * fstride:i * fstride + rows * ls].reshape(rows, ls)
        fr[:, :cols] = util.natural_image(rows, cols, seed=900 + i, integer=True)
    res = []
    for padded in (0, 1):
        wm = _mk(wmb, rows, cols, W)
        wm.set_option(wmb.OPT_PADDED_UPLOAD, padded)
        out = np.zeros((n, rows, cols), np.uint8)
        sc = np.zeros(2 * n, np.float32)
        ctx = wmb.VideoProcessingContext(wm, rows, cols, interval, linesize=ls, frame_stride=fstride, frames_on_device=False)
        wmb.process_frames(ctx, wmb.VIDEO_EMBED_VERIFY, buf.ctypes.data, out.ctypes.data, first, n, sc)
        res.append((out, sc))
        wm.close()
    assert np.array_equal(res[0][0], res[1][0]) and np.array_equal(res[0][1], res[1][1], equal_nan=True)
    out, sc = res[1]
    for i in range(n):
        src = buf[i * fstride:i * fstride + rows * ls].reshape(rows, ls)[:, :cols]
        if (first + i) % interval:
            assert np.array_equal(out[i], src) and np.isnan(sc[i]) and np.isnan(sc[n + i])
            continue
        st, o, a = oracle.embed_frame_u8(src, W, 40.0, oracle.ME)
        assert st == 0 and abs(sc[i] - a) <= 1e-3 * a
        assert int(np.abs(out[i].astype(np.int32) - o.astype(np.int32)).max()) <= 1


@pytest.mark.parametrize("fused", [0, 1])
def test_singular_and_zero_mask_through_the_replayed_sync_path(wmb, oracle, fused):
    """The error behaviour of Watermark.cpp:164-165,246-247 on the path the reference's loops_for_test protocol takes: repeated synchronous
    calls (captured / replayed graph, result delivered by the op's last CTA into mapped pinned memory), then a solvable image on the same
    context, then the constant one again — no stale scalar may leak from one op into the next."""
    rows, cols = 96, 160
    const = np.full((rows, cols), 100.0, np.float32)
    nat = util.natural_image(rows, cols, seed=21)
    W = util.normal_w(rows, cols)
    wm = _mk(wmb, rows, cols, W)
    wm.set_option(wmb.OPT_FUSED_SINGLE, fused)
    dc = wmb.DeviceArray.from_numpy(wm, const, wmb.COL_MAJOR)
    dn = wmb.DeviceArray.from_numpy(wm, nat, wmb.COL_MAJOR)
    out = wmb.DeviceArray(wm, rows, cols, wmb.COL_MAJOR, wmb.F32)
    on = oracle.embed(nat, W, 40.0, wmb.ME)
    for rep in range(2):
        for _ in range(4):
            _, a, st = wm.makeWatermark(dc, dc, wmb.ME, out=out)
            assert st == wmb.SINGULAR and np.isnan(a) and np.array_equal(out.numpy(), const)
        for _ in range(4):
            _, a, st = wm.makeWatermark(dc, dc, wmb.NVF, out=out)
            assert st == wmb.ZERO_MASK and np.isinf(a) and np.array_equal(out.numpy(), const)
        for mask in (wmb.ME, wmb.NVF):
            for _ in range(4):
                corr, st = wm.detectWatermark(dc, mask)
                assert st == wmb.SINGULAR and corr == 0.0
        for _ in range(4):
            _, a, st = wm.makeWatermark(dn, dn, wmb.ME, out=out)
            assert st == 0 and abs(a - on["a"]) <= 1e-3 * on["a"]
        z = out.numpy().copy()
        od = oracle.detect(z, W, wmb.ME)
        for _ in range(4):
            corr, st = wm.detectWatermark(out, wmb.ME)
            assert st == 0 and abs(corr - od["corr"]) <= 1e-3 * abs(od["corr"])
    wm.close()


# ---------------------------------------------------------------------------------------------------
# u8 frames: stats / apply on 128-thread CTAs (8 lines per thread, the default) == on 256-thread CTAs
# ---------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("rows,cols,ls", [(64, 64, 64), (96, 160, 192), (270, 480, 480), (130, 264, 272), (1080, 1920, 1920)])
def test_narrow_u8_kernels_equal_wide_ones(wmb, oracle, rows, cols, ls):
    """WM_OPT_NARROW_U8 (stats, apply, detector): same arithmetic per pixel; a thread's f32 partials of a tile now cover 32 pixels instead of
    16, so the strength and the correlation may move in their last bits (<= 1e-6 relative) — the watermarked bytes must still agree to 1 LSB (observed: identical) and the
    frames must match the oracle like the wide kernels do."""
    n = 5
    W = util.normal_w(rows, cols)
    frames = np.zeros((n, rows, ls), np.uint8)
    for i in range(n):
        frames[i, :, :cols] = util.natural_image(rows, cols, seed=700 + i, integer=True)
        frames[i, :, cols:] = 9
    frames[3, :, :cols] = 77  # unsolvable system: frame copied through by the apply kernel
    res = {}
    for narrow in (1, 0):
        wm = _mk(wmb, rows, cols, W)
        wm.set_option(wmb.OPT_NARROW_U8, narrow)
        L = wmb.lib()
        dfr = L.wm_dev_alloc(wm._h, frames.nbytes)
        dout = L.wm_dev_alloc(wm._h, n * rows * cols)
        L.wm_dev_upload(wm._h, dfr, frames.ctypes.data, frames.nbytes)
        ctx = wmb.VideoProcessingContext(wm, rows, cols, 1, linesize=ls, frames_on_device=True)
        a = np.zeros(n, np.float32)
        wmb.process_frames(ctx, wmb.VIDEO_EMBED, dfr, dout, 0, n, a)
        out = np.zeros((n, rows, cols), np.uint8)
        L.wm_dev_download(wm._h, out.ctypes.data, dout, out.nbytes)
        # the detector on the watermarked frames (identical bytes for both variants, asserted below)
        cdet = np.zeros(n, np.float32)
        ctx_o = wmb.VideoProcessingContext(wm, rows, cols, 1, linesize=cols, frames_on_device=True)
        wmb.process_frames(ctx_o, wmb.VIDEO_DETECT, dout, None, 0, n, cdet)
        res[narrow] = (a, out, cdet)
        L.wm_dev_free(wm._h, dfr)
        L.wm_dev_free(wm._h, dout)
        wm.close()
    (a1, o1, c1), (a0, o0, c0) = res[1], res[0]
    assert c1[3] == 0.0 and c0[3] == 0.0  # unsolvable frame: correlation 0 (Watermark.cpp:246-247)
    okc = c0 != 0.0
    assert np.max(np.abs(c1[okc] - c0[okc]) / np.abs(c0[okc])) <= 2e-6
    st2, oc = oracle.detect_frame_u8(o1[2], W, oracle.ME)
    assert abs(c1[2] - oc) <= 1e-3 * abs(oc)
    assert np.isnan(a1[3]) and np.isnan(a0[3]) and np.array_equal(o1[3], frames[3, :, :cols])
    ok = ~np.isnan(a0)
    assert np.max(np.abs(a1[ok] - a0[ok]) / a0[ok]) <= 1e-6
    assert int(np.abs(o1.astype(np.int32) - o0.astype(np.int32)).max()) <= 1
    st, oo, oa = oracle.embed_frame_u8(frames[2], W, 40.0, oracle.ME, width=cols)
    assert abs(a1[2] - oa) <= 1e-3 * oa and int(np.abs(o1[2].astype(np.int32) - oo.astype(np.int32)).max()) <= 1
    report("narrow vs wide u8 kernels %dx%d ls=%d: a rel diff %.2g, differing bytes %d, corr rel diff %.2g" % (
        rows, cols, ls, float(np.max(np.abs(a1[ok] - a0[ok]) / a0[ok])), int(np.count_nonzero(o1 != o0)),
        float(np.max(np.abs(c1[okc] - c0[okc]) / np.abs(c0[okc])))))


@pytest.mark.parametrize("p", [3, 5])
@pytest.mark.parametrize("rows,cols", [(96, 160), (270, 480)])
def test_u8_images_both_masks_narrow_kernels(wmb, oracle, rows, cols, p):
    """u8 images through the batch API with BOTH masks (the video driver only ever uses ME): the 128-thread u8 kernels with the fused 3 x 3
    NVF mask (p = 3) and with the precomputed mask planes of a larger window (p = 5: stats / detector in the 128-thread form, apply in the
    256-thread one on the same grid), against the oracle's u8 frame path."""
    B = 3
    W = util.normal_w(rows, cols)
    o = oracle.opts(p=p)
    ys = np.stack([util.natural_image(rows, cols, seed=950 + b, integer=True) for b in range(B)])
    for narrow in (1, 0):
        wm = wmb.Watermark(rows, cols, W, p, 40.0)
        wm.set_option(wmb.OPT_NARROW_U8, narrow)
        L = wmb.lib()
        din = L.wm_dev_alloc(wm._h, ys.nbytes)
        dout = L.wm_dev_alloc(wm._h, ys.nbytes)
        L.wm_dev_upload(wm._h, din, ys.ctypes.data, ys.nbytes)
        di = wmb.image_desc(din, rows, cols, wmb.ROW_MAJOR, wmb.U8)
        do = wmb.image_desc(dout, rows, cols, wmb.ROW_MAJOR, wmb.U8)
        for mask in (wmb.NVF, wmb.ME):
            ab, cb = np.zeros(B, np.float32), np.zeros(B, np.float32)
            wm.embed_batch(0, di, di, do, rows * cols, rows * cols, rows * cols, B, mask, ab)
            wm.detect_batch(0, do, rows * cols, B, mask, cb)
            wm.sync(0)
            out = np.zeros_like(ys)
            L.wm_dev_download(wm._h, out.ctypes.data, dout, out.nbytes)
            for b in range(B):
                st, oo, oa = oracle.embed_frame_u8(ys[b], W, 40.0, mask, o=o)
                st2, oc = oracle.detect_frame_u8(oo, W, mask, o=o)
                assert st == 0 and abs(ab[b] - oa) <= 1e-3 * oa, (narrow, mask, b, ab[b], oa)
                assert int(np.abs(out[b].astype(np.int32) - oo.astype(np.int32)).max()) <= 1
                assert abs(cb[b] - oc) <= 1e-3 * abs(oc), (narrow, mask, b, cb[b], oc)
        L.wm_dev_free(wm._h, din)
        L.wm_dev_free(wm._h, dout)
        wm.close()
