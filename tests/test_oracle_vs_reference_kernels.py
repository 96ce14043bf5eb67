"""Pins the C restatement (oracle/wm_oracle.c) against the reference's own kernel source executed on the CPU
(oracle/clshim -> oracle/_ref/libref_kernels_*.so, built from /root/reference/Watermark_GPU/kernels/*.hpp).
The reference ships no golden vectors, so this is the pin for the three OpenCL kernels; the
ArrayFire calls between them stay a restatement of documented semantics."""
import os

import numpy as np
import pytest

import util

SHAPES = [(64, 64), (48, 80), (67, 131), (130, 70), (512, 512)]


@pytest.fixture(scope="module")
def ref():
    from oracle import ref_kernels
    if not ref_kernels.available():
        # the only external pin of the three kernels must not disappear silently: oracle/_ref is built by `make -C oracle` where
        # /root/reference exists and travels (prebuilt) to boxes without it
        if os.environ.get("WM_ALLOW_MISSING_REF") == "1":
            pytest.skip("oracle/_ref not built and WM_ALLOW_MISSING_REF=1")
        pytest.fail("oracle/_ref is missing: run `make -C oracle` where /root/reference is present (or set WM_ALLOW_MISSING_REF=1 to skip the pin)")
    return ref_kernels


@pytest.mark.parametrize("rows,cols", SHAPES)
def test_nvf_kernel(oracle, ref, rows, cols):
    img = util.natural_image(rows, cols, seed=1) if rows != 512 else util.load_512_gray(oracle)
    assert np.array_equal(oracle.nvf(img, oracle.opts(contract=0)), ref.nvf(img, fma=False))
    # with contraction allowed (-cl-mad-enable, main.cpp:106) WHICH mul+add pairs fuse is the compiler's choice, and
    # the naive variance amplifies it (8.6e-3 max abs between contraction choices): bounded, not equal
    assert np.abs(oracle.nvf(img, oracle.opts(contract=1)) - ref.nvf(img, fma=True)).max() <= 1e-2


@pytest.mark.parametrize("p", [5, 7, 9])
@pytest.mark.parametrize("rows,cols", [(64, 64), (67, 131), (130, 70)])
def test_nvf_kernel_larger_windows(oracle, ref, rows, cols, p):
    """The class accepts p in {3, 5, 7, 9} (Watermark.cpp:24); the nvf kernel is generic in p (kernels/nvf.hpp:14-17)."""
    img = util.natural_image(rows, cols, seed=10 + p)
    assert np.array_equal(oracle.nvf(img, oracle.opts(contract=0, p=p)), ref.nvf(img, fma=False, p=p))
    assert np.abs(oracle.nvf(img, oracle.opts(contract=1, p=p)) - ref.nvf(img, fma=True, p=p)).max() <= 2e-2


@pytest.mark.parametrize("rows,cols", SHAPES)
def test_scaled_neighbors_kernel(oracle, ref, rows, cols):
    img = util.natural_image(rows, cols, seed=2) if rows != 512 else util.load_512_gray(oracle)
    coef = np.array([-0.21, 0.47, -0.18, 0.49, 0.52, -0.2, 0.45, -0.24], np.float32)
    assert np.array_equal(oracle.scaled_neighbors(img, coef, oracle.opts(contract=0)), ref.scaled_neighbors(img, coef, fma=False))
    assert np.array_equal(oracle.scaled_neighbors(img, coef, oracle.opts(contract=1)), ref.scaled_neighbors(img, coef, fma=True))


@pytest.mark.parametrize("rows,cols", SHAPES)
def test_me_kernel(oracle, ref, rows, cols):
    """fp16-rounded products, 64-wide sequential f32 group sums, padded columns contributing 0."""
    img = util.natural_image(rows, cols, seed=3) if rows != 512 else util.load_512_gray(oracle)
    Rx, rx = oracle.rx(img, oracle.FAITHFUL)
    for fma in (False, True):  # products are stored to fp16 before any add: contraction cannot change them
        rRx, rrx = ref.rx(img, fma=fma)
        assert np.array_equal(Rx, rRx)
        assert np.array_equal(rx, rrx)
    # integer-valued frames
    y = util.natural_image(rows, cols, seed=4, integer=True).astype(np.float32)
    Rx, rx = oracle.rx(y, oracle.FAITHFUL)
    rRx, rrx = ref.rx(y)
    assert np.array_equal(Rx, rRx) and np.array_equal(rx, rrx)
    assert np.array_equal(Rx, Rx.T)


def test_me_kernel_partial_layout(ref):
    """moddims(RxPartial, 64, n) / moddims(rxPartial, 8, n) (Watermark.cpp:148-149): one 8x8 block per work-group."""
    img = util.natural_image(16, 100, seed=5)
    Rxp, rxp = ref.me_partials(img)
    assert Rxp.shape == (16 * 128 // 64, 64) and rxp.shape == (16 * 128 // 64, 8)
    blk = Rxp[1].reshape(8, 8)  # second group of row 0: columns 64..99 valid, 100..127 padded
    assert np.array_equal(blk, blk.T) and blk[0, 0] > 0
