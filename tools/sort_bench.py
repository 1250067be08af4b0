"""Microbenchmark of the radix sort (K2) through the C ABI: n pairs, `bits` significant key bits.

    python tools/sort_bench.py [n] [bits] [u32|u64] [reps]
"""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from octreelib_b200 import _native as N
from octreelib_b200.forest import TorchAllocator

n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 100_000_000
bits = int(sys.argv[2]) if len(sys.argv) > 2 else 19
kind = sys.argv[3] if len(sys.argv) > 3 else "u64"
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 5
dev = torch.device("cuda", 0)
lib = N.lib()
alloc = TorchAllocator(dev)
g = torch.Generator(device=dev).manual_seed(0)
src = torch.randint(0, 1 << bits, (n,), dtype=torch.int64 if kind == "u64" else torch.int32, device=dev, generator=g)
vals0 = torch.arange(n, dtype=torch.int32, device=dev)
fn = lib.ol_sort_pairs_u64 if kind == "u64" else lib.ol_sort_pairs_u32
stream = torch.cuda.current_stream(dev)
ksz = 8 if kind == "u64" else 4
passes = (bits + 7) // 8
best = 1e9
for r in range(reps):
    k = src.clone()
    v = vals0.clone()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    N.check(fn(C.c_void_p(stream.cuda_stream), C.c_void_p(k.data_ptr()), C.c_void_p(v.data_ptr()), n, 0, bits, alloc.alloc_cb,
               alloc.free_cb, None))
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    best = min(best, ms)
    algo = n * (ksz + passes * 2 * (ksz + 4))
    print(f"rep {r}: {ms:.3f} ms, {n / ms / 1e6:.2f} Gkeys/s, algorithmic {algo / ms / 1e6:.0f} GB/s ({passes} passes)", flush=True)
assert bool((k[1:] >= k[:-1]).all()), "not sorted"
print(f"best {best:.3f} ms")
