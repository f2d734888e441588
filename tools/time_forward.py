"""Small synthesis calls as a user makes them: synthesize_flat (device-resident) and GriffinLimVocoder.forward on one
utterance (numpy's global generator continued on the device): host enqueue time, time until done, single-call latency."""
import importlib, os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
pkg = importlib.import_module(bench.PKG)
voc = pkg.GriffinLimVocoder(24000, 1200, 300, 2048, 80, 20, 8000, torch.hann_window, spec_bwd_max_iter=64).cuda()
for T in [int(a) for a in sys.argv[1:]] or [100, 500]:
    x = torch.from_numpy(bench.synth_logmel_np(T, 1)).cuda()
    ph = ((torch.rand(T, 1025, device="cuda") * 2 - 1) * np.pi).contiguous()
    tw = time.perf_counter()
    while time.perf_counter() - tw < 1.5:
        for _ in range(20): y = voc.synthesize_flat(x, [T], ph)
        torch.cuda.synchronize()
    import pynvml
    pynvml.nvmlInit(); h = pynvml.nvmlDeviceGetHandleByIndex(0)
    print("   SM clock after warm-up", pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM), "MHz")
    t0 = time.perf_counter()
    for _ in range(20): y = voc.synthesize_flat(x, [T], ph)
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    print(f"T={T}: host enqueue {1e3 * (t1 - t0) / 20:.3f} ms/call, until done {1e3 * (t2 - t0) / 20:.3f} ms/call", flush=True)
    # one call alone, timed with events
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record(); y = voc.synthesize_flat(x, [T], ph); e1.record(); torch.cuda.synchronize()
    print(f"   single call, events: {e0.elapsed_time(e1):.3f} ms")
    # the reference's call: forward() on one utterance (numpy's global RNG continued on the device)
    np.random.seed(0)
    for _ in range(5): y = voc(x)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(20): y = voc(x)
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    print(f"   forward(): host {1e3 * (t1 - t0) / 20:.3f} ms/call, until done {1e3 * (t2 - t0) / 20:.3f} ms/call -> {(T - 1) * 300 / 24000 / ((t2 - t0) / 20):.0f} audio-s/s", flush=True)
    t0 = time.perf_counter(); y = voc(x); torch.cuda.synchronize(); t1 = time.perf_counter()
    print(f"   forward(), one call with sync: {1e3 * (t1 - t0):.3f} ms")
