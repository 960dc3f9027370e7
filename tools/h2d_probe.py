"""Why is the end-to-end upload slow?  Pinned host -> device bandwidth measured several ways:
python tools/h2d_probe.py   (prints one JSON line per experiment)"""
import ctypes, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch

dev = torch.device("cuda", 0)
torch.zeros(1, device=dev)
rt = ctypes.CDLL("libcudart.so.12")
rt.cudaHostAlloc.argtypes = [ctypes.POINTER(ctypes.c_void_p), ctypes.c_size_t, ctypes.c_uint]
rt.cudaMemcpyAsync.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int, ctypes.c_void_p]


def timed(fn, n=10):
    fn(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n):
        fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n


def report(what, nbytes, ms):
    print(json.dumps({"what": what, "MB": round(nbytes / 1e6, 2), "ms": round(ms, 4), "GB/s": round(nbytes / ms / 1e6, 2)}), flush=True)


for mb in (1, 8, 45, 75, 256):
    n = mb << 20
    d = torch.empty(n, dtype=torch.uint8, device=dev)
    # (1) torch pin_memory of a numpy-backed tensor (what bench.py does)
    h1 = torch.from_numpy(np.random.randint(0, 255, n, dtype=np.uint8)).pin_memory()
    report("torch.from_numpy().pin_memory() -> copy_", n, timed(lambda: d.copy_(h1, non_blocking=True)))
    # (2) torch.empty(pin_memory=True)
    h2 = torch.empty(n, dtype=torch.uint8, pin_memory=True); h2.copy_(h1)
    report("torch.empty(pin_memory=True) -> copy_", n, timed(lambda: d.copy_(h2, non_blocking=True)))
    # (3) cudaHostAlloc directly, flags default / write-combined
    for flags, name in ((0, "cudaHostAlloc default"), (4, "cudaHostAlloc write-combined")):
        p = ctypes.c_void_p()
        assert rt.cudaHostAlloc(ctypes.byref(p), n, flags) == 0
        ctypes.memmove(p.value, h1.data_ptr(), n)
        st = torch.cuda.current_stream().cuda_stream
        report(name + " -> cudaMemcpyAsync", n,
               timed(lambda: rt.cudaMemcpyAsync(ctypes.c_void_p(d.data_ptr()), p, n, 1, ctypes.c_void_p(st))))
        rt.cudaFreeHost(p)
    # (4) D2H for comparison
    report("D2H copy_", n, timed(lambda: h2.copy_(d, non_blocking=True)))
# NUMA / affinity facts
try:
    print(json.dumps({"cpu_count": os.cpu_count(), "affinity": len(os.sched_getaffinity(0)),
                      "numa_nodes": sorted(x for x in os.listdir("/sys/devices/system/node") if x.startswith("node"))}))
    import subprocess
    print(subprocess.run(["nvidia-smi", "topo", "-m"], capture_output=True, text=True).stdout[:1500])
    print(subprocess.run(["nvidia-smi", "--query-gpu=pcie.link.gen.current,pcie.link.width.current,pcie.link.gen.max,pcie.link.width.max",
                          "--format=csv"], capture_output=True, text=True).stdout)
except Exception as e:
    print("topology probe failed:", e)
