"""Raw pinned-memory PCIe bandwidth of the box (H2D, D2H, both at once): the ceiling of bench.py's e2e number.
Single GPU: `python tools/pcie_bw.py`.  All GPUs at once (aggregate host <-> device ceiling of the box):
`python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 tools/pcie_bw.py`."""
import os
import time

import torch


def main():
    rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    dist = None
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("gloo")
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    n = 256 << 20
    h_in = torch.empty(n, dtype=torch.uint8).pin_memory()
    h_out = torch.empty(n, dtype=torch.uint8).pin_memory()
    d_a = torch.empty(n, dtype=torch.uint8, device=dev)
    d_b = torch.empty(n, dtype=torch.uint8, device=dev)
    s1, s2 = torch.cuda.Stream(dev), torch.cuda.Stream(dev)

    def run(h2d, d2h, reps=8):
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        t = time.perf_counter()
        for _ in range(reps):
            if h2d:
                with torch.cuda.stream(s1):
                    d_a.copy_(h_in, non_blocking=True)
            if d2h:
                with torch.cuda.stream(s2):
                    h_out.copy_(d_b, non_blocking=True)
        torch.cuda.synchronize()
        gbs = reps * n / (time.perf_counter() - t) / 1e9
        if dist is not None:
            v = torch.tensor([gbs], dtype=torch.float64)
            dist.all_reduce(v)
            gbs = float(v.item())
        return gbs

    run(True, True, 2)
    res = [("H2D alone", run(True, False)), ("D2H alone", run(False, True)), ("both at once (each direction)", run(True, True))]
    if rank == 0:
        for name, v in res:
            print("%-30s %.1f GB/s%s" % (name, v, " summed over %d GPUs" % world if world > 1 else ""))
    if world == 1:
        fb = 1080 * 1920 * 4  # chunked like the e2e path: 8.3 MB frames
        nf = n // fb
        torch.cuda.synchronize()
        t = time.perf_counter()
        for r in range(4):
            for i in range(nf):
                with torch.cuda.stream(s2):
                    h_out[i * fb:(i + 1) * fb].copy_(d_b[i * fb:(i + 1) * fb], non_blocking=True)
        torch.cuda.synchronize()
        print("D2H in 8.3 MB frames           %.1f GB/s" % (4 * nf * fb / (time.perf_counter() - t) / 1e9))
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
