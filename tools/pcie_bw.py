"""Raw pinned-memory PCIe bandwidth of the box (H2D, D2H, both at once): the ceiling of bench.py's e2e number."""
import time

import torch


def main():
    dev = torch.device("cuda", 0)
    n = 256 << 20
    h_in = torch.empty(n, dtype=torch.uint8).pin_memory()
    h_out = torch.empty(n, dtype=torch.uint8).pin_memory()
    d_a = torch.empty(n, dtype=torch.uint8, device=dev)
    d_b = torch.empty(n, dtype=torch.uint8, device=dev)
    s1, s2 = torch.cuda.Stream(dev), torch.cuda.Stream(dev)

    def run(h2d, d2h, reps=8):
        torch.cuda.synchronize()
        t = time.perf_counter()
        for _ in range(reps):
            if h2d:
                with torch.cuda.stream(s1):
                    d_a.copy_(h_in, non_blocking=True)
            if d2h:
                with torch.cuda.stream(s2):
                    h_out.copy_(d_b, non_blocking=True)
        torch.cuda.synchronize()
        return reps * n / (time.perf_counter() - t) / 1e9

    run(True, True, 2)
    print("H2D alone      %.1f GB/s" % run(True, False))
    print("D2H alone      %.1f GB/s" % run(False, True))
    print("both at once   %.1f GB/s each direction" % run(True, True))
    # chunked like the e2e path: 8.3 MB frames
    fb = 1080 * 1920 * 4
    nf = n // fb
    torch.cuda.synchronize()
    t = time.perf_counter()
    for r in range(4):
        for i in range(nf):
            with torch.cuda.stream(s2):
                h_out[i * fb:(i + 1) * fb].copy_(d_b[i * fb:(i + 1) * fb], non_blocking=True)
    torch.cuda.synchronize()
    print("D2H in 8.3 MB frames %.1f GB/s" % (4 * nf * fb / (time.perf_counter() - t) / 1e9))


if __name__ == "__main__":
    main()
