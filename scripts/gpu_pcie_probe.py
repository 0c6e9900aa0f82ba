"""What the host link of this box sustains for 128 MiB pinned transfers: H2D alone, D2H alone, both at once (two streams).
The denominator for the NTT's end-to-end figure (uzkge_cuda_ntt_fr_batch moves 128 MiB each way per 2^22 transform)."""
import json

import torch

n = 128 << 20
h_in = torch.empty(n, dtype=torch.uint8).pin_memory()
h_out = torch.empty(n, dtype=torch.uint8).pin_memory()
d_a = torch.empty(n, dtype=torch.uint8, device="cuda")
d_b = torch.empty(n, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def timed(fn, reps=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    torch.cuda.synchronize()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def h2d():
    with torch.cuda.stream(s1):
        d_a.copy_(h_in, non_blocking=True)


def d2h():
    with torch.cuda.stream(s2):
        h_out.copy_(d_b, non_blocking=True)


def both():
    h2d()
    d2h()


out = {}
for name, fn in (("h2d_alone", h2d), ("d2h_alone", d2h), ("both_at_once", both)):
    ms = timed(fn)
    out[name] = {"ms_per_128MiB": ms, "GB_per_s_each_way": n / (ms * 1e-3) / 1e9}
print(json.dumps(out))
