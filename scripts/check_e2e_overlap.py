"""End-to-end inference step time with / without the H2D prefetcher and the D2H sink, eager and graph replay: found that a small
H2D copy in front of a replay queues behind the next batch's volume upload on the copy engine (hence coma_upload_small)."""
import os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from coma_unet_b200 import DevicePrefetcher, HostSink
from coma_unet_b200.graph import GraphedInference
dev = torch.device("cuda", 0)
shape = (128, 128, 128)
model = bench.build_model(dev, shape); model.eval(); model.set_training(False)
mri, tau, roi, covars, dicts = bench.make_batch(8, 1234, shape, device=dev)
h_mri, h_tau, h_roi, h_cov, _ = bench.make_batch(8, 1234, shape, pin=True)
g = GraphedInference(model)
def eager(m, r, c):
    with torch.no_grad():
        return model(m, c, roi_pred_dicts=dicts, sample_roi_mask=r)
def graphed(m, r, c):
    return g(m, c, dicts, r)
sink = HostSink((8, 1, *shape), torch.float32, dev)
def run(fn, n, use_sink=True, use_prefetch=True):
    def hb():
        for _ in range(n): yield (h_mri, h_roi)
    host = 0.0
    it = DevicePrefetcher(hb(), dev) if use_prefetch else ((mri, roi) for _ in range(n))
    for m, r in it:
        t0 = time.perf_counter()
        out = fn(m, r, h_cov)
        host += time.perf_counter() - t0
        if use_sink: sink.put(out)
    sink.wait()
    return host
for name, fn in (("eager", eager), ("graph", graphed)):
    for use_sink, use_pf in ((True, True), (False, True), (True, False), (False, False)):
        run(fn, 6, use_sink, use_pf); torch.cuda.synchronize()
        t0 = time.perf_counter(); host = run(fn, 60, use_sink, use_pf); torch.cuda.synchronize(); dt = time.perf_counter() - t0
        print(f"{name:6s} sink={use_sink} prefetch={use_pf}: {dt / 60 * 1e3:.2f} ms/step, host time in the step call {host / 60 * 1e3:.2f} ms", flush=True)
