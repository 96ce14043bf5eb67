"""CPU-side tests (-m "not gpu"): the oracle against its golden vectors and the surveyor's independent numbers,
the host logic, and that the C-ABI library loads, exports every declared symbol and fails loudly without a GPU."""
import ctypes
import json
import os
import re

import numpy as np
import pytest

import util

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = json.load(open(os.path.join(util.GOLDEN, "golden_512.json")))


# ---------------------------------------------------------------------------------------------------
# oracle vs golden vectors (tests/golden/make_golden.py) and vs SURVEY.md §8c (independent numpy restatement)
# ---------------------------------------------------------------------------------------------------
def test_oracle_strength(oracle):
    assert oracle.strength(40.0) == pytest.approx(2.55, rel=1e-6)  # Watermark.cpp:22
    assert oracle.strength(40.0) == GOLD["inputs"]["strength_psnr40"]


@pytest.mark.parametrize("mode", ["faithful", "exact", "strict_f32"])
def test_oracle_matches_golden(oracle, mode):
    opt = {"faithful": oracle.FAITHFUL, "exact": oracle.EXACT, "strict_f32": oracle.STRICT_F32}[mode]
    img = util.load_512_gray(oracle)
    W = util.load_w512()
    g = GOLD["modes"][mode]
    pe = oracle.pred_error_mask(img, opt)
    assert np.array_equal(pe["Rx"], np.array(g["Rx"]))
    assert np.array_equal(pe["rx"], np.array(g["rx"]))
    assert np.array_equal(pe["coef"], np.array(g["coef"], np.float32))
    assert np.array_equal(pe["e"][100:104, 200:204], np.array(g["e_crop"], np.float32))
    nv = oracle.nvf(img, opt)
    assert np.array_equal(nv[100:104, 200:204], np.array(g["nvf_crop"], np.float32))
    for mask, mn in ((oracle.NVF, "nvf"), (oracle.ME, "me")):
        e = oracle.embed(img, W, 40.0, mask, o=opt)
        assert e["a"] == g[mn]["a"]
        d = oracle.detect(e["out"], W, mask, o=opt)
        assert d["corr"] == g[mn]["corr_marked"]
        assert abs(g[mn]["psnr"] - 40.0) < 0.02  # MSE == strength^2 by construction (Watermark.cpp:170)


def test_oracle_matches_survey_numbers(oracle):
    """SURVEY.md §8c table (numpy restatement by the surveyor, f64 sums) — an independent implementation."""
    f, e = GOLD["modes"]["faithful"], GOLD["modes"]["exact"]
    assert f["nvf"]["a"] == pytest.approx(2.852794, rel=2e-6)
    assert f["me"]["a"] == pytest.approx(34.90134, rel=2e-6)
    assert e["me"]["a"] == pytest.approx(34.90300, rel=2e-6)
    assert f["nvf"]["corr_marked"] == pytest.approx(0.585582, abs=5e-5)
    assert e["nvf"]["corr_marked"] == pytest.approx(0.585916, abs=5e-5)
    assert f["me"]["corr_marked"] == pytest.approx(0.737925, abs=5e-5)
    assert e["me"]["corr_marked"] == pytest.approx(0.737665, abs=5e-5)
    assert abs(f["me"]["corr_clean"]) < 0.005 and abs(f["nvf"]["corr_clean"]) < 0.005


def test_oracle_video_frame_golden(oracle):
    img = util.load_512_gray(oracle)
    W = util.load_w512()
    y = np.rint(img).astype(np.uint8)
    st, out, a = oracle.embed_frame_u8(y, W, 40.0, oracle.ME)
    st2, corr = oracle.detect_frame_u8(out, W, oracle.ME)
    assert st == 0 and st2 == 0
    assert a == GOLD["video_u8"]["a"] and corr == GOLD["video_u8"]["corr"]
    # row padding is skipped (main.cpp:348-353)
    pad = np.full((512, 576), 200, np.uint8)
    pad[:, :512] = y
    st, out2, a2 = oracle.embed_frame_u8(pad, W, 40.0, oracle.ME, width=512)
    assert a2 == a and np.array_equal(out, out2)


def test_oracle_layout_equivalence(oracle):
    """Processing the transposed image gives the transposed result with permuted coefficients —
    the identity the CUDA path's column-major mode relies on."""
    img = util.natural_image(96, 72, seed=2, integer=True).astype(np.float32)  # integer pixels: sums are exact
    a = oracle.pred_error_mask(img, oracle.EXACT)
    b = oracle.pred_error_mask(np.ascontiguousarray(img.T), oracle.EXACT)
    perm = [0, 3, 5, 1, 6, 2, 4, 7]
    assert np.array_equal(a["Rx"], b["Rx"][np.ix_(perm, perm)])
    assert np.array_equal(a["rx"], b["rx"][perm])
    assert np.allclose(a["coef"], b["coef"][perm], rtol=1e-5, atol=1e-7)
    assert np.allclose(a["e"], b["e"].T, atol=1e-3)


def test_oracle_lag_identity_integer(oracle):
    """The lag-symmetric regrouping used by k_sweep, checked exactly on integer pixels with numpy."""
    rng = np.random.default_rng(0)
    H, Wd = 23, 37
    X = rng.integers(0, 256, (H, Wd)).astype(np.int64)
    Rx, rx = oracle.rx(X.astype(np.float32), oracle.EXACT)
    off = [(-1, -1), (-1, 0), (-1, 1), (0, -1), (0, 1), (1, -1), (1, 0), (1, 1)]

    def Xc(r, c):
        return X[np.clip(r, 0, H - 1), np.clip(c, 0, Wd - 1)]

    rr, cc = np.meshgrid(np.arange(H), np.arange(Wd), indexing="ij")
    core = (rr >= 1) & (rr <= H - 2) & (cc >= 1) & (cc <= Wd - 2)
    for i in range(8):
        for j in range(i, 8):
            d = (off[j][0] - off[i][0], off[j][1] - off[i][1])
            s_core = int(np.sum(X[core] * Xc(rr + d[0], cc + d[1])[core]))
            qi_r, qi_c = rr + off[i][0], cc + off[i][1]
            q_in_core = (qi_r >= 1) & (qi_r <= H - 2) & (qi_c >= 1) & (qi_c <= Wd - 2)
            s_frame = int(np.sum((Xc(qi_r, qi_c) * Xc(rr + off[j][0], cc + off[j][1]))[~q_in_core]))
            assert s_core + s_frame == int(Rx[i, j])


def test_oracle_singular_and_degenerate(oracle):
    img = np.full((64, 64), 100.0, np.float32)
    W = util.normal_w(64, 64)
    e = oracle.embed(img, W, 40.0, oracle.ME)
    assert e["status"] == 1 and np.array_equal(e["out"], img)
    d = oracle.detect(img, W, oracle.ME)
    assert d["status"] == 1 and d["corr"] == 0.0
    e = oracle.embed(img, W, 40.0, oracle.NVF)
    assert e["status"] == 2 and np.array_equal(e["out"], img)


# ---------------------------------------------------------------------------------------------------
# the C ABI: loads, exports what include/wm_b200.h declares, fails loudly without a device
# ---------------------------------------------------------------------------------------------------
def _declared_symbols():
    h = open(os.path.join(ROOT, "include", "wm_b200.h")).read()
    h = re.sub(r"/\*.*?\*/", "", h, flags=re.S)
    return sorted(set(re.findall(r"\b(wm_[a-z0-9_]+)\s*\(", h)))


def test_cabi_exports_every_declared_symbol(wmb):
    L = wmb.lib()
    decl = _declared_symbols()
    assert len(decl) >= 30
    for s in decl:
        assert hasattr(L, s), "libwm_b200.so does not export " + s
    assert sorted(wmb.EXPORTS) == decl
    assert b"sm_100a" in L.wm_version()


def test_cabi_struct_layout(wmb):
    assert ctypes.sizeof(wmb.wm_image) == 56
    assert ctypes.sizeof(wmb.wm_video_ctx) == 40


def test_no_cpu_fallback(wmb):
    """Without a GPU the product path must fail loudly (no oracle / CPU route behind the ABI)."""
    if wmb.device_count() > 0:
        pytest.skip("a GPU is present")
    with pytest.raises(wmb.WatermarkError) as e:
        wmb.Watermark(64, 64, util.normal_w(64, 64), 3, 40.0)
    assert e.value.code == -7
    src = open(os.path.join(ROOT, "watermarking-gpu_b200", "__init__.py")).read()
    code = re.sub(r'""".*?"""', "", src, flags=re.S)
    assert "oracle" not in code
    for f in ("wm_api.cu", "wm_kernels.cuh"):
        assert "oracle" not in open(os.path.join(ROOT, "watermarking-gpu_b200", "csrc", f)).read()


def test_argument_validation_without_device(wmb):
    L = wmb.lib()
    h = ctypes.c_void_p()
    w = util.normal_w(64, 64)
    fp = w.ctypes.data_as(ctypes.POINTER(ctypes.c_float))
    assert L.wm_create(ctypes.byref(h), 64, 64, fp, 4, 40.0, 0, None) == -1     # Watermark.cpp:24-25
    assert b"Wrong p parameter: 4" in L.wm_last_error(None)
    assert L.wm_create(ctypes.byref(h), 64, 64, fp, 3, -1.0, 0, None) == -5     # main.cpp:96
    assert L.wm_create(ctypes.byref(h), 2, 64, fp, 3, 40.0, 0, None) == -4


def test_shard_frames(wmb):
    for n in (1, 7, 64, 513):
        for world in (1, 2, 4, 8):
            seen = []
            for r in range(world):
                first, cnt = wmb.shard_frames(n, r, world)
                seen += list(range(first, first + cnt))
            assert seen == list(range(n))  # contiguous, disjoint, complete, in rank order


# ---------------------------------------------------------------------------------------------------
# multi-process plumbing (gloo, world_size 2): shard by frame index, no data-path collective, gather scalars
# ---------------------------------------------------------------------------------------------------
def _worker(rank, world, port, q):
    import importlib
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    wmb = importlib.import_module("watermarking-gpu_b200")
    n, interval = 13, 3
    first, cnt = wmb.shard_frames(n, rank, world)
    # each rank "processes" its frames: the gate uses the GLOBAL index (main.cpp:346)
    local = torch.tensor([float(i) if i % interval == 0 else float("nan") for i in range(first, first + cnt)])
    sizes = [wmb.shard_frames(n, r, world)[1] for r in range(world)]
    bufs = [torch.zeros(s) for s in sizes]
    dist.all_gather(bufs, local) if len(set(sizes)) == 1 else [
        dist.broadcast(bufs[r] if r != rank else local, src=r) for r in range(world)]
    bufs[rank] = local
    t = torch.tensor([1.0 + rank])
    dist.all_reduce(t, op=dist.ReduceOp.MAX)  # max-over-ranks timing
    if rank == 0:
        q.put((torch.cat(bufs).tolist(), float(t.item())))
    dist.destroy_process_group()


def test_two_rank_sharding_gloo():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 500)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    vals, tmax = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert tmax == 2.0
    exp = [float(i) if i % 3 == 0 else float("nan") for i in range(13)]
    assert all((np.isnan(a) and np.isnan(b)) or a == b for a, b in zip(vals, exp))


# ---------------------------------------------------------------------------------------------------
# watermark generator: reproduces the stream of the committed .dat files
# ---------------------------------------------------------------------------------------------------
def test_common_random_matrix_reproduces_committed_file(wmb, tmp_path):
    import subprocess
    exe = wmb.build().build_generator()
    out = tmp_path / "w.dat"
    r = subprocess.run([exe, "512", "512", "28390211", str(out)], capture_output=True, text=True)
    assert r.returncode == 0 and "Successfully wrote 262144 random floats" in r.stdout
    a = np.fromfile(out, np.float32)
    b = util.load_w512().ravel()
    assert a.size == b.size
    ulp = np.abs(a.view(np.int32).astype(np.int64) - b.view(np.int32).astype(np.int64))
    assert ulp.max() <= 2 and np.abs(a - b).max() <= 5e-7          # MSVC evaluates the polar transform in double
    # all committed files are prefixes of one stream: a smaller request is a prefix of a larger one
    out2 = tmp_path / "w2.dat"
    subprocess.run([exe, "64", "64", "28390211", str(out2)], check=True, capture_output=True)
    assert np.array_equal(np.fromfile(out2, np.float32), a[:4096])
    # argument validation (CommonRandomMatrix/main.cpp:17-32)
    assert subprocess.run([exe, "0", "64", "1", str(out2)], capture_output=True).returncode != 0
    assert subprocess.run([exe, "64"], capture_output=True).returncode != 0
    # the file loads through the library's own size check (Watermark.cpp:70-71) when a GPU is present
    assert abs(float(a.mean())) < 0.01 and abs(float(a.std()) - 1.0) < 0.01


def test_reference_arm_prints_the_contract_line():
    """bench.py --impl reference: the reference's path on the host cores (the CPU restatement; no GPU, no product code), one JSON line
    with the keys the driver compares against the GPU arm's."""
    import json
    import subprocess
    import sys
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                       capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "frames/s" and d["higher_is_better"] is True and d["value"] > 0
    assert d["config"]["workload"] == "video4k" and d["config"]["rows"] == 2160 and d["config"]["cols"] == 3840
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["gpu_launches"] == 0
