import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, pynvml
import bench
pynvml.nvmlInit()
h = pynvml.nvmlDeviceGetHandleByIndex(0)
dev = torch.device("cuda", 0)
w = bench.WORKLOADS["c4_street_100M"]
clouds, numbers, P, total = bench.make_workload("c4_street_100M", 0, 1, dev, 1.0)
for _ in range(3):
    bench.run_step(clouds, numbers, P, w, 1)
def sample():
    t0 = time.perf_counter()
    a = pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)
    t1 = time.perf_counter()
    b = pynvml.nvmlDeviceGetCurrentClocksEventReasons(h)
    t2 = time.perf_counter()
    c = pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM)
    t3 = time.perf_counter()
    return a, b, c, 1e3*(t1-t0), 1e3*(t2-t1), 1e3*(t3-t2)
for it in range(4):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    bench.run_step(clouds, numbers, P, w, 1)
    torch.cuda.synchronize(); t1 = time.perf_counter()
    s = sample()
    # and a sample while the GPU is busy: enqueue a long kernel first
    x = torch.empty(1 << 28, device=dev); x.normal_()
    s2 = sample()
    torch.cuda.synchronize()
    print(f"step {1e3*(t1-t0):.1f} ms; idle sample {s}; busy sample {s2}", flush=True)
