#!/usr/bin/env python
"""bench.py — embed+detect FPS of the watermark hot path on B200 (BASELINE.json metric), one JSON line.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--workload image1080p|video4k|image4k|image8k|batch256|image512]
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...
  python bench.py --impl reference ...      # the reference path on the host CPU cores (oracle port, all threads)

A "step" is one pass of the hot path over one batch of synthetic frames that are resident in HBM before the
timed region: every frame goes through NVF embed, ME embed, NVF detect and ME detect (image workloads; the
reference's testForImage protocol, main.cpp:167-223) or ME embed + ME detect (video workload, main.cpp:343-410).
`value` = frames/s over all ranks (device-timed, max over ranks); `e2e` = the same metric through the
host-buffer C-ABI calls with pinned HOST buffers (H2D + D2H inside the timed region).
The batch is larger than L2 (126 MB), so every step streams its inputs from HBM.
"""
import argparse
import importlib
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

WORKLOADS = {
    # name: (rows, cols, frames per step, kind, dtype).  Frames per step divide the resident CTA counts of a launch (3 x 148
    # for sweep / stats / apply, 2 x 148 for the detector), so that every image of a batch gets the same number of CTAs
    "image512": (512, 512, 256, "image", "f32"),
    "image1080p": (1080, 1920, 148, "image", "f32"),
    "image4k": (2160, 3840, 37, "image", "f32"),
    "image8k": (4320, 7680, 4, "image", "f32"),
    "batch256": (256, 256, 4096, "image", "f32"),
    "video4k": (2160, 3840, 37, "video", "u8"),
}
# algorithmic (compulsory) bytes per pixel and kernel: every distinct operand read once, every output written
# once (SURVEY.md §8d / DESIGN.md): f32 image, f32 W, f32 out; u8 frames: 1-byte pixels
ALG_BYTES = {
    "f32": {"rx_sweep": 4, "me_stats": 8, "nvf_stats": 8, "embed_apply": 12, "detect_apply": 8},
    "u8": {"rx_sweep": 1, "me_stats": 5, "nvf_stats": 5, "embed_apply": 6, "detect_apply": 5},
}
PAIR_BYTES = {"f32": 20, "u8": 11}  # embed + detect per pixel


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.proc = [], None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "50"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), line.strip()))

    def stop(self, t0, t1):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], None, set()
        for ts, line in self.rows:
            f = [x.strip() for x in line.split(",")]
            if len(f) < 7:
                continue
            try:
                mx = float(f[1])
                if t0 - 0.05 <= ts <= t1 + 0.15:
                    sm.append(float(f[0]))
                    for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[3:7]):
                        if v.lower().startswith("active"):
                            reasons.add(name)
            except ValueError:
                continue
        if not sm:  # region shorter than the sampling period: take whatever was seen
            for ts, line in self.rows:
                try:
                    sm.append(float(line.split(",")[0]))
                except ValueError:
                    pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm)}


def make_inputs(rows, cols, n, dtype):
    import util
    base = util.natural_image(rows, cols, seed=1)
    rng = np.random.default_rng(7)
    out = np.empty((n, rows, cols), np.uint8 if dtype == "u8" else np.float32)
    for i in range(n):  # distinct natural-statistics frames: cyclic shift + light noise
        f = np.roll(base, (3 * i + 1, 5 * i + 2), (0, 1)) + rng.uniform(-2, 2, base.shape).astype(np.float32)
        f = np.clip(f, 0, 255)
        out[i] = np.rint(f) if dtype == "u8" else f
    W = util.normal_w(rows, cols, seed=28390211)
    return out, W


# ------------------------------------------------------------------------------------------------
# reference arm / cpu_baseline: the oracle port on the host cores
# ------------------------------------------------------------------------------------------------
def cpu_frame(oracle, img, W, kind):
    """one frame through the same ops as the GPU step"""
    if kind == "video":
        st, out, a = oracle.embed_frame_u8(img, W, 40.0, oracle.ME)
        oracle.detect_frame_u8(out, W, oracle.ME)
        return
    for mask in (oracle.NVF, oracle.ME):
        o = oracle.embed(img, W, 40.0, mask)
        oracle.detect(o["out"], W, mask)


def cpu_baseline(oracle, frames, W, kind, budget_s=12.0):
    t = time.perf_counter()
    cpu_frame(oracle, frames[0], W, kind)  # warm-up + calibration
    t1 = time.perf_counter() - t
    n = int(max(2, min(len(frames), round(budget_s / max(t1, 1e-3)))))
    t = time.perf_counter()
    for i in range(n):
        cpu_frame(oracle, frames[i % len(frames)], W, kind)
    dt = time.perf_counter() - t
    return n / dt, n


def use_all_host_threads():
    """torchrun exports OMP_NUM_THREADS=1; the CPU arm is meant to use every host core (libgomp reads this at load)."""
    os.environ["OMP_NUM_THREADS"] = str(os.cpu_count() or 1)
    try:  # undo bind_to_gpu_numa_node: the CPU arm uses every core of the box
        os.sched_setaffinity(0, range(os.cpu_count() or 1))
    except OSError:
        pass


def run_reference(args, wl):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    use_all_host_threads()
    from oracle import oracle
    rows, cols, nfr, kind, dtype = WORKLOADS[wl]
    frames, W = make_inputs(rows, cols, min(nfr, 4), dtype)
    # bounded sample per step so the whole run ends within minutes
    t = time.perf_counter()
    cpu_frame(oracle, frames[0], W, kind)
    t1 = time.perf_counter() - t
    per_step = int(max(1, min(8, round(3.0 / max(t1, 1e-3)))))
    for _ in range(args.warmup):
        cpu_frame(oracle, frames[0], W, kind)
    t = time.perf_counter()
    for s in range(args.steps):
        for i in range(per_step):
            cpu_frame(oracle, frames[(s * per_step + i) % len(frames)], W, kind)
    dt = time.perf_counter() - t
    fps = args.steps * per_step / dt
    line = {
        "impl": "reference", "metric": "embed+detect FPS (NVF & PE masks)" if kind == "image" else "embed+detect FPS (PE mask, u8 video frames)",
        "value": fps, "unit": "frames/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32" if dtype == "f32" else "u8->f32", "data": "synthetic",
        "config": {"workload": wl, "rows": rows, "cols": cols, "frames_per_step": per_step, "p": 3, "psnr": 40.0,
                   "ops_per_frame": "NVF embed, ME embed, NVF detect, ME detect" if kind == "image" else "ME embed, ME detect"},
        "cpu_baseline": {"value": fps, "unit": "frames/s", "cores": oracle.num_threads(), "kind": "port",
                         "sample": "%d frame(s) per step x %d steps of %s on the host CPU (OpenMP oracle restating "
                                   "Watermark.cpp; ArrayFire/OpenCL are not installable offline)" % (per_step, args.steps, wl)},
        "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))
    return 0


# ------------------------------------------------------------------------------------------------
# B200 arm
# ------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="image1080p", choices=sorted(WORKLOADS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-secondary", action="store_true", help="skip the video4k record a default (image1080p) run appends")
    ap.add_argument("--no-sync-proto", action="store_true", help="skip the single-image synchronous-call measurement (keeps ncu launch lists to the timed region)")
    ap.add_argument("--exact", action="store_true", help="f32 products in Rx/rx instead of the reference's fp16 rounding")
    ap.add_argument("--frames", type=int, default=0, help="frames per step (0 = the workload's own count)")
    ap.add_argument("--no-tma", action="store_true", help="plain register-prefetched tile loaders instead of TMA (A/B)")
    ap.add_argument("--two-streams", action="store_true", help="A/B: NVF ops and ME ops of a step on two slots (kernel tails overlap)")
    ap.add_argument("--split-cost", type=int, default=None, help="WM_OPT_SPLIT_COST (A/B: -1 = never partition an unbalanced batch)")
    ap.add_argument("--fhadd", action="store_true", help="sum the rounded Rx/rx products with the FHADD chain instead of HMMA (A/B)")
    ap.add_argument("--chunk", type=int, default=0, help="frames per API call (0 = the whole batch in one call)")
    ap.add_argument("--slots", type=int, default=2, help="pipeline slots (streams) used round-robin when --chunk is set")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    wl = args.workload
    if args.impl == "reference":
        return run_reference(args, wl)

    import torch
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    dist = None
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    bind_to_gpu_numa_node(torch, local_rank)
    line = run_workload(args, wl, torch, dist, dev, rank, local_rank, world, args.frames)
    # the other headline of BASELINE.json's metric (configs[2]: 4K u8 video, ME embed + detect per frame) rides along on a default
    # run as a compact secondary record; `value`, `roofline`, `e2e` above stay those of the primary workload
    if args.workload == "image1080p" and not args.no_secondary:
        sec = run_workload(args, "video4k", torch, dist, dev, rank, local_rank, world, 0, secondary=True)
        if line is not None and sec is not None:
            line["secondary"] = {"video4k": {k: sec[k] for k in ("metric", "value", "unit", "ms_per_step", "dtype", "config", "step_frac_of_peak", "e2e", "gpu_launches")}}
            line["secondary"]["video4k"]["kernels"] = [{"kernel": k["kernel"], "avg_ms": k["avg_ms"], "achieved_gbs": k["achieved_gbs"]} for k in sec["kernels"]]
    if line is not None:
        print(json.dumps(line))
    if dist is not None:
        dist.destroy_process_group()
    return 0


NUMA_NOTE = None


def bind_to_gpu_numa_node(torch, index):
    """Pin this rank's host threads to the CPUs next to its GPU BEFORE any pinned buffer is allocated (first touch puts the pages
    on that NUMA node), so the e2e copies of several ranks do not all cross one socket's memory controllers / the inter-socket
    link.  Best effort: NVML may be missing or the box may have one node."""
    global NUMA_NOTE
    if os.environ.get("WM_BENCH_NO_AFFINITY"):
        NUMA_NOTE = "affinity off (WM_BENCH_NO_AFFINITY)"
        return
    try:
        import pynvml
        pynvml.nvmlInit()
        pr = torch.cuda.get_device_properties(index)
        bus = "%08x:%02x:%02x.0" % (pr.pci_domain_id, pr.pci_bus_id, pr.pci_device_id)
        h = pynvml.nvmlDeviceGetHandleByPciBusId(bus.encode() if bytes is str else bus)
        pynvml.nvmlDeviceSetCpuAffinity(h)
        cpus = sorted(os.sched_getaffinity(0))
        NUMA_NOTE = "host threads bound to the %d CPUs next to GPU %s (%d..%d)" % (len(cpus), bus, cpus[0], cpus[-1])
    except Exception as e:  # noqa: BLE001
        NUMA_NOTE = "no NUMA binding (%s)" % (str(e)[:80],)


def run_workload(args, wl, torch, dist, dev, rank, local_rank, world, frames_override, secondary=False):
    """One workload through warm-up, the timed region, e2e and (primary only) the CPU baseline; returns the JSON record on rank 0."""
    pkg = importlib.import_module("watermarking-gpu_b200")
    C = pkg.C

    rows, cols, nfr, kind, dtype = WORKLOADS[wl]
    if frames_override:
        nfr = frames_override
    npx = rows * cols
    frames_np, W = make_inputs(rows, cols, nfr, dtype)
    stream = torch.cuda.Stream(device=dev)
    wm = pkg.Watermark(rows, cols, W, 3, 40.0, device=local_rank, stream=stream.cuda_stream)
    if args.exact:
        wm.set_option(pkg.OPT_FP16_PRODUCTS, 0)
    if args.fhadd:
        wm.set_option(pkg.OPT_MMA_ACCUM, 0)
    if args.no_tma:
        wm.set_option(pkg.OPT_USE_TMA, 0)
    if args.split_cost is not None:
        wm.set_option(pkg.OPT_SPLIT_COST, args.split_cost)
    tdt = torch.uint8 if dtype == "u8" else torch.float32
    dt_code = pkg.U8 if dtype == "u8" else pkg.F32
    # image workloads: ArrayFire layout (column-major); video: row-major Y planes
    layout = pkg.ROW_MAJOR if kind == "video" else pkg.COL_MAJOR
    mem = frames_np if layout == pkg.ROW_MAJOR else np.ascontiguousarray(frames_np.transpose(0, 2, 1))
    d_in = torch.from_numpy(mem).to(dev)
    d_out = [torch.empty_like(d_in) for _ in range(2)]  # NVF-marked, ME-marked
    a_host = [np.zeros(nfr, np.float32) for _ in range(2)]
    c_host = [np.zeros(nfr, np.float32) for _ in range(2)]
    st_host = np.zeros(nfr, np.int32)
    di = pkg.image_desc(d_in.data_ptr(), rows, cols, layout, dt_code)
    do = [pkg.image_desc(t.data_ptr(), rows, cols, layout, dt_code) for t in d_out]

    vctx_e = pkg.VideoProcessingContext(wm, rows, cols, 1, linesize=cols, frames_on_device=True)

    def step():
        if kind == "video":
            pkg.process_frames(vctx_e, pkg.VIDEO_EMBED, d_in.data_ptr(), d_out[1].data_ptr(), 0, nfr, a_host[1])
            pkg.process_frames(vctx_e, pkg.VIDEO_DETECT, d_out[1].data_ptr(), None, 0, nfr, c_host[1])
            return
        if not args.chunk:
            for k, mask in enumerate((pkg.NVF, pkg.ME)):
                sl = k if args.two_streams else 0  # A/B: the NVF chain and the ME chain are independent
                wm.embed_batch(sl, di, di, do[k], npx, npx, npx, nfr, mask, a_host[k], st_host)
                wm.detect_batch(sl, do[k], npx, nfr, mask, c_host[k], st_host)
            wm.sync(-1 if args.two_streams else 0)
            return
        # chunked: each call covers `chunk` frames so that one op's passes (sweep -> stats -> apply) find the frames
        # in L2; calls go round-robin over `slots` streams so launch gaps and per-image solves overlap
        esz = d_in.element_size()
        for phase in range(4):
            k, mask = (0, pkg.NVF) if phase % 2 == 0 else (1, pkg.ME)
            for ci, c0 in enumerate(range(0, nfr, args.chunk)):
                n = min(args.chunk, nfr - c0)
                slot = ci % args.slots
                off = c0 * npx * esz
                dic = pkg.image_desc(d_in.data_ptr() + off, rows, cols, layout, dt_code)
                doc = pkg.image_desc(d_out[k].data_ptr() + off, rows, cols, layout, dt_code)
                if phase < 2:
                    wm.embed_batch(slot, dic, dic, doc, npx, npx, npx, n, mask, a_host[k][c0:c0 + n])
                else:
                    wm.detect_batch(slot, doc, npx, n, mask, c_host[k][c0:c0 + n])
        wm.sync(-1)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize(dev)

    with torch.cuda.stream(stream):
        sampler = ClockSampler(local_rank) if rank == 0 else None
        t_warm = time.time()
        for _ in range(args.warmup):
            step()
        while time.time() - t_warm < 0.5:  # keep the GPU under the same load until the sampler has a few readings
            step()
        barrier()
        wm.set_option(pkg.OPT_KERNEL_TIMING, 1)
        wm.kernel_times(reset=True)
        l0 = wm.launch_count
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.time()
        e0.record(stream)
        for _ in range(args.steps):
            step()
        e1.record(stream)
        barrier()
        t1 = time.time()
        ms = e0.elapsed_time(e1)
        launches = wm.launch_count - l0
        ktimes = wm.kernel_times(reset=True)
        k_steps = args.steps  # steps the per-kernel times cover
        wm.set_option(pkg.OPT_KERNEL_TIMING, 0)
        if kind == "video":
            # the driver keeps 4 slots in flight, so the events above bracket kernels that share the GPU; per-kernel
            # durations for the roofline come from one more pass with the slots serialised (wm option 4)
            wm.set_option(4, 1)
            wm.set_option(pkg.OPT_KERNEL_TIMING, 1)
            k_steps = max(2, args.steps // 4)
            for _ in range(k_steps):
                step()
            ktimes = wm.kernel_times(reset=True)
            wm.set_option(pkg.OPT_KERNEL_TIMING, 0)
            wm.set_option(4, 0)
    clocks = sampler.stop(t_warm + 0.2, t1) if sampler else None
    if dist is not None:
        t = torch.tensor([ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    fps = world * nfr * args.steps / (ms * 1e-3)

    # ---- e2e: frames start in pinned HOST memory; per frame the timed region holds the H2D copy of the frame, the
    # hot path through the C ABI (device-pointer entry points on the ctx's slot streams, 4 frames in flight) and the
    # D2H copy of every result (watermarked frames + scalars).  Detection runs on the watermarked frame where the
    # embed left it (on the device), as in the reference's testForImage / video loop.
    e2e = None
    if not args.no_e2e:
        n_e2e = min(nfr, 32)
        NS = wm.num_slots
        ext = [torch.cuda.ExternalStream(wm.stream(sl), device=dev) for sl in range(NS)]
        pin_in = torch.from_numpy(mem[:n_e2e].copy()).pin_memory()
        nout = 1 if kind == "video" else 2
        pin_out = [torch.empty_like(pin_in).pin_memory() for _ in range(nout)]
        s_in = [torch.empty_like(d_in[0]) for _ in range(NS)]
        s_out = [[torch.empty_like(d_in[0]) for _ in range(nout)] for _ in range(NS)]
        a_e = [np.zeros(n_e2e, np.float32) for _ in range(nout)]
        c_e = [np.zeros(n_e2e, np.float32) for _ in range(nout)]
        fb = pin_in[0].numel() * pin_in.element_size()
        masks = (pkg.ME,) if kind == "video" else (pkg.NVF, pkg.ME)

        def e2e_frame(i):
            sl = i % NS
            with torch.cuda.stream(ext[sl]):
                s_in[sl].copy_(pin_in[i], non_blocking=True)
                din = pkg.image_desc(s_in[sl].data_ptr(), rows, cols, layout, dt_code)
                for k2, mask in enumerate(masks):
                    dout = pkg.image_desc(s_out[sl][k2].data_ptr(), rows, cols, layout, dt_code)
                    wm.embed_batch(sl, din, din, dout, 0, 0, 0, 1, mask, a_e[k2][i:i + 1])
                    pin_out[k2][i].copy_(s_out[sl][k2], non_blocking=True)
                for k2, mask in enumerate(masks):
                    dout = pkg.image_desc(s_out[sl][k2].data_ptr(), rows, cols, layout, dt_code)
                    wm.detect_batch(sl, dout, 0, 1, mask, c_e[k2][i:i + 1])

        for i in range(min(2 * NS, n_e2e)):
            e2e_frame(i)
        wm.sync(-1)
        barrier()
        reps = max(1, args.steps // 4)
        tt = time.perf_counter()
        for r in range(reps):
            for i in range(n_e2e):
                e2e_frame(i)
        wm.sync(-1)
        torch.cuda.synchronize(dev)
        e2e_s = time.perf_counter() - tt
        if dist is not None:
            t = torch.tensor([e2e_s], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            e2e_s = float(t.item())
        # the same frames through the e2e path must give the same scalars as the resident path
        e2e_ok = bool(np.allclose(a_e[-1][:n_e2e], a_host[1][:n_e2e], rtol=1e-6) and np.allclose(c_e[-1][:n_e2e], c_host[1][:n_e2e], rtol=1e-5, atol=1e-7))
        e2e = {"value": world * reps * n_e2e / e2e_s, "unit": "frames/s",
               "h2d_bytes_per_step": int(fb * n_e2e), "d2h_bytes_per_step": int(nout * fb * n_e2e + 2 * nout * 4 * n_e2e),
               "frames_per_step": n_e2e, "frames_in_flight": NS, "matches_resident_path": e2e_ok,
               "h2d_gbs": world * reps * n_e2e / e2e_s * fb / 1e9, "d2h_gbs": world * reps * n_e2e / e2e_s * nout * fb / 1e9,
               "host_affinity": NUMA_NOTE,
               "pcie_note": "tools/pcie_bw.py on this pool's boxes (profiles/r1_pcie_bw.txt): one GPU copies 55 GB/s one way, 46 GB/s each way when both "
                            "directions run; all 8 GPUs together only 118 GB/s D2H / 77 GB/s each way, which caps e2e at N = 8",
               "api": "wm_embed_batch / wm_detect_batch on wm_get_stream() slots; pinned host frames in, watermarked frames + scalars out"}

    # the reference's literal protocol for this config (main.cpp:167-223): ONE image, synchronous calls, mean over loops
    sync_proto = None
    if kind == "image" and rank == 0 and not args.no_sync_proto and not secondary:
        one_in = pkg.image_desc(d_in.data_ptr(), rows, cols, layout, dt_code)
        one_out = pkg.image_desc(d_out[0].data_ptr(), rows, cols, layout, dt_code)
        av, cv = C.c_float(0), C.c_float(0)
        L_ = pkg.lib()
        loops = 200

        def four_ops():
            for mask in (pkg.NVF, pkg.ME):
                L_.wm_embed(wm._h, C.byref(one_in), C.byref(one_in), C.byref(one_out), mask, C.byref(av))
                L_.wm_detect(wm._h, C.byref(one_out), mask, C.byref(cv))

        for _ in range(3):
            four_ops()
        tt = time.perf_counter()
        for _ in range(loops):
            four_ops()
        dt_s = time.perf_counter() - tt
        sync_proto = {"value": loops / dt_s, "unit": "frames/s", "us_per_frame": 1e6 * dt_s / loops, "loops": loops,
                      "what": "one resident image, synchronous wm_embed / wm_detect calls (4 per frame), CUDA-graph replay"}

    # multi-GPU: the only data that crosses ranks — per-frame scalars, gathered after the timed region
    gathered = None
    if dist is not None:
        mine = torch.from_numpy(np.stack([a_host[1], c_host[1]])).to(dev)
        parts = [torch.empty_like(mine) for _ in range(world)]
        dist.all_gather(parts, mine)
        gathered = torch.stack(parts).cpu().numpy()  # [rank, (a, corr), frame]
    if rank != 0:
        return None

    # ---- roofline of the dominant kernel (largest share of device time in the timed region) ----
    peak, peak_src = peaks()
    kern = []
    for name, (n, tot_ms) in ktimes.items():
        if n == 0:
            continue
        per_launch_px = npx * nfr * k_steps * {"rx_sweep": (3 if kind == "image" else 2), "embed_apply": 2 if kind == "image" else 1, "detect_apply": 2 if kind == "image" else 1}.get(name, 1) / n  # pixels one launch covers
        alg = ALG_BYTES[dtype][name] * per_launch_px
        avg_ms = tot_ms / n
        kern.append({"kernel": name, "launches": n, "avg_ms": avg_ms, "total_ms": tot_ms,
                     "alg_bytes_per_launch": alg, "achieved_gbs": alg / (avg_ms * 1e-3) / 1e9})
    kern.sort(key=lambda k: -k["total_ms"])
    dom = kern[0] if kern else None
    roof = None
    if dom:
        roof = {"bound": "hbm", "kernel": dom["kernel"], "achieved": dom["achieved_gbs"], "peak": peak, "unit": "GB/s",
                "frac": dom["achieved_gbs"] / peak, "traffic": None, "peak_source": peak_src,
                "alg_bytes_per_launch": dom["alg_bytes_per_launch"], "avg_launch_ms": dom["avg_ms"]}
        tr = os.path.join(ROOT, "profiles", "traffic.json")  # DRAM bytes per frame from the ncu --set full capture
        if os.path.exists(tr):
            try:
                per_frame = json.load(open(tr)).get(wl, {}).get("per_frame", {}).get(dom["kernel"])
                if per_frame is not None:
                    roof["traffic"] = per_frame * dom["alg_bytes_per_launch"] / (ALG_BYTES[dtype][dom["kernel"]] * npx)
                    roof["traffic_source"] = "profiles/traffic.json (ncu dram__bytes_read.sum + dram__bytes_write.sum per frame x frames per launch)"
            except Exception:
                pass
    # whole-step effective bandwidth on the compulsory bytes of embed+detect pairs
    pairs = 2 if kind == "image" else 1
    step_bytes = PAIR_BYTES[dtype] * npx * nfr * pairs
    cb = None
    parity = None
    if not args.no_cpu_baseline and world == 1 and not secondary:  # rank 0 at N = 1 only
        use_all_host_threads()
        from oracle import oracle
        v, n = cpu_baseline(oracle, frames_np, W, kind)
        # parity gate in the same job: frame 0 through the oracle against the scalars the timed region produced
        if kind == "image":
            chk = {}
            for k2, mask in enumerate((oracle.NVF, oracle.ME)):
                o = oracle.embed(frames_np[0], W, 40.0, mask)
                od = oracle.detect(o["out"], W, mask)
                chk["a_%s_rel" % ("nvf", "me")[k2]] = abs(float(a_host[k2][0]) - o["a"]) / abs(o["a"])
                chk["corr_%s_rel" % ("nvf", "me")[k2]] = abs(float(c_host[k2][0]) - od["corr"]) / abs(od["corr"])
        else:
            st_, out_, a_ = oracle.embed_frame_u8(frames_np[0], W, 40.0, oracle.ME)
            corr_ = oracle.detect_frame_u8(out_, W, oracle.ME)[1]
            chk = {"a_me_rel": abs(float(a_host[1][0]) - a_) / abs(a_), "corr_me_rel": abs(float(c_host[1][0]) - corr_) / abs(corr_)}
        parity = dict(chk, tolerance=1e-3, ok=bool(all(v_ <= 1e-3 for v_ in chk.values())), what="frame 0: GPU scalars of the timed region vs the CPU oracle")
        cb = {"value": v, "unit": "frames/s", "cores": oracle.num_threads(), "kind": "port",
              "sample": "%d frames of %s, same ops per frame, OpenMP oracle (restates Watermark.cpp; the reference's "
                        "ArrayFire/OpenCL stack cannot be built offline)" % (n, wl)}
    line = {
        "metric": "embed+detect FPS (NVF & PE masks)" if kind == "image" else "embed+detect FPS (PE mask, u8 video frames)",
        "value": fps, "unit": "frames/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms / args.steps, "wall_ms_per_step": (t1 - t0) * 1e3 / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32" if dtype == "f32" else "u8->f32", "data": "synthetic",
        "config": {"workload": wl, "rows": rows, "cols": cols, "frames_per_step": nfr, "p": 3, "psnr": 40.0,
                   "layout": "col-major (ArrayFire)" if layout == pkg.COL_MAJOR else "row-major Y plane",
                   "ops_per_frame": "NVF embed, ME embed, NVF detect, ME detect" if kind == "image" else "ME embed, ME detect",
                   "fp16_products": not args.exact, "frames_per_call": args.chunk or nfr, "slots": args.slots if args.chunk else 1,
                   "l2": "inputs larger than L2: %.0f MB of frames + W per step" % ((d_in.numel() * d_in.element_size() + W.nbytes) / 1e6),
                   "sharding": "frames sharded by rank, no collective on the data path"},
        "step_effective_gbs": step_bytes / (ms / args.steps * 1e-3) / 1e9,
        "step_frac_of_peak": step_bytes / (ms / args.steps * 1e-3) / 1e9 / peak,
        "roofline": roof, "kernels": kern, "cpu_baseline": cb, "e2e": e2e, "gpu_launches": int(launches),
        "sync_single_image": sync_proto, "parity": parity,
        "clocks": clocks,
        "results": {"a_nvf": float(a_host[0][0]), "a_me": float(a_host[1][0]), "corr_nvf": float(c_host[0][0]),
                    "corr_me": float(c_host[1][0]),
                    "scalars_gathered": None if gathered is None else
                    {"ranks": int(gathered.shape[0]), "frames_per_rank": int(gathered.shape[2]),
                     "corr_me_mean": float(gathered[:, 1].mean()), "a_me_mean": float(gathered[:, 0].mean())}},
    }
    return line


if __name__ == "__main__":
    sys.exit(main())
