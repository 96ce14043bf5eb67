"""Runs the C++ façade program (tests/cpp/test_facade.cpp: Watermark.hpp + videoprocessingcontext.hpp used the way
the reference's main.cpp uses its class) on the GPU and compares what it prints with the CPU oracle."""
import os
import subprocess

import numpy as np
import pytest

import util

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_cpp_facade_against_oracle(wmb, oracle, tmp_path):
    b = wmb.build()
    exe = b.build_facade_test()
    assert exe and os.path.exists(exe)
    rows, cols, nframes, ls = 120, 200, 5, 224
    img = util.natural_image(rows, cols, seed=21)
    W = util.normal_w(rows, cols)
    frames = np.zeros((nframes, rows, ls), np.uint8)
    for i in range(nframes):
        frames[i, :, :cols] = util.natural_image(rows, cols, seed=400 + i, integer=True)
        frames[i, :, cols:] = 77
    img.tofile(tmp_path / "gray.f32")
    W.tofile(tmp_path / "w.dat")
    frames.tofile(tmp_path / "frames.u8")
    r = subprocess.run([exe, str(tmp_path / "gray.f32"), str(tmp_path / "w.dat"), str(rows), str(cols),
                        str(tmp_path / "frames.u8"), str(nframes), str(ls)], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    got = dict(line.split(None, 1) for line in r.stdout.strip().splitlines())
    assert got["bad_p"].strip() == "1" and got["bad_w"].strip() == "1"
    onv = oracle.embed(img, W, 40.0, oracle.NVF)
    ome = oracle.embed(img, W, 40.0, oracle.ME)
    assert float(got["a_nvf"]) == pytest.approx(onv["a"], rel=1e-3)
    assert float(got["a_me"]) == pytest.approx(ome["a"], rel=1e-3)
    assert float(got["corr_nvf"]) == pytest.approx(oracle.detect(onv["out"], W, oracle.NVF)["corr"], rel=1e-3)
    cme = oracle.detect(ome["out"], W, oracle.ME)["corr"]
    assert float(got["corr_me"]) == pytest.approx(cme, rel=1e-3)
    assert float(got["corr_me_copy"]) == float(got["corr_me"]) == float(got["corr_me_assigned"])
    # every u8 pixel within 1 LSB => the checksum differs by at most rows*cols
    assert abs(int(got["sum_u8_me"]) - int(ome["out"].astype(np.uint8).astype(np.int64).sum())) <= rows * cols * 0.01
    assert int(got["frames"]) == nframes
    for i in range(nframes):
        if i % 2 == 0:
            st, oo, oa = oracle.embed_frame_u8(frames[i], W, 40.0, oracle.ME, width=cols)
            st2, oc = oracle.detect_frame_u8(oo, W, oracle.ME)
            assert abs(int(got["frame_sum_%d" % i]) - int(oo.astype(np.int64).sum())) <= rows * cols * 0.01
            assert float(got["frame_corr_%d" % i]) == pytest.approx(oc, rel=1e-3)
        else:  # gated off: Y plane passes through with the padding dropped, no detection
            assert int(got["frame_sum_%d" % i]) == int(frames[i, :, :cols].astype(np.int64).sum())
            assert float(got["frame_corr_%d" % i]) == 0.0
    # processFramesInMemory over two contexts (chunks of the global index, one host thread each) == the per-frame functions
    assert got["batched_equal_pixels"].strip() == "1"
    for i in range(nframes):
        assert float(got["batched_corr_%d" % i]) == float(got["frame_corr_%d" % i])
