"""PCIe probe for the e2e arm: pinned H2D alone, D2H alone, both at once (two streams), with 1 and 16 chunks.
Tooling only (uses torch for streams/pinned memory); prints GB/s."""
import time
import torch

N = 256 * 1024 * 1024  # bytes per direction
h_in = torch.empty(N, dtype=torch.uint8, pin_memory=True)
h_out = torch.empty(N, dtype=torch.uint8, pin_memory=True)
d_in = torch.empty(N, dtype=torch.uint8, device="cuda")
d_out = torch.empty(N, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def run(h2d, d2h, chunks, reps=5):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        c = N // chunks
        for k in range(chunks):
            if h2d:
                with torch.cuda.stream(s1):
                    d_in[k * c:(k + 1) * c].copy_(h_in[k * c:(k + 1) * c], non_blocking=True)
            if d2h:
                with torch.cuda.stream(s2):
                    h_out[k * c:(k + 1) * c].copy_(d_out[k * c:(k + 1) * c], non_blocking=True)
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / reps
    return N / dt / 1e9


for chunks in (1, 16, 256):
    a = run(True, False, chunks)
    b = run(False, True, chunks)
    c = run(True, True, chunks)
    print("chunks %4d  H2D %.1f GB/s  D2H %.1f GB/s  both: %.1f GB/s per direction (%.1f aggregate)" % (chunks, a, b, c, 2 * c))
