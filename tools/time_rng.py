"""Device time of s2st_phase_from_mt19937 (the sequential MT19937 recurrence + the parallel phase kernel) and the host
cost of the shim around it (draw_initial_phase_device with / without the state hand-back)."""
import importlib, os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
pkg = importlib.import_module(bench.PKG)
vm = importlib.import_module(pkg.__name__ + ".vocoder")
L = pkg._lib
dev = torch.device("cuda", 0)
torch.zeros(1, device=dev)
for T in (100, 500, 2000):
    n = 1025 * T
    st = np.random.get_state()
    key = torch.from_numpy(np.asarray(st[1], np.uint32).view(np.int32).copy()).to(dev)
    words = torch.empty(2 * n, dtype=torch.int32, device=dev); key_out = torch.empty(624, dtype=torch.int32, device=dev)
    phase = torch.empty(T, 1025, device=dev)
    def run():
        L.check(L.load().s2st_phase_from_mt19937(1, 1025, T, L.ptr(key), int(st[2]), L.ptr(words), L.ptr(phase), L.ptr(key_out), L.stream_ptr(dev)), "x")
    for _ in range(3): run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): run()
    e1.record(); torch.cuda.synchronize()
    print(f"T={T}: generator + phase kernels {e0.elapsed_time(e1) / 10:.3f} ms for {2 * n} words")
    # host-side cost of the shim
    np.random.seed(0)
    t0 = time.perf_counter()
    for _ in range(20):
        ph, fin = vm.draw_initial_phase_device((1025, T), dev)
    t1 = time.perf_counter()
    for _ in range(20):
        ph, fin = vm.draw_initial_phase_device((1025, T), dev); fin()
    t2 = time.perf_counter()
    print(f"      shim without finish {1e3 * (t1 - t0) / 20:.3f} ms/call, with finish (sync + set_state) {1e3 * (t2 - t1) / 20:.3f} ms/call")
