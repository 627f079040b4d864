import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
dev = torch.device("cuda", 0); shape = (128,)*3
model = bench.build_model(dev, shape); model.eval(); model.set_training(False)
mri, tau, roi, covars, dicts = bench.make_batch(8, 1234, shape, device=dev)
def run():
    with torch.no_grad(): return model(mri, covars, roi_pred_dicts=dicts, sample_roi_mask=roi)
def timeit(n=30):
    for _ in range(5): run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): out = run()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n, out
base_ms, base = timeit()
for fl in (True,):
    model.fusion_layer.fuse_prologue = fl; model.deep_modulator_3c.fuse_prologue = fl
    ms, out = timeit()
    print("fuse_prologue", fl, f"{ms:.3f} ms vs {base_ms:.3f} ms; max diff vs unfused {float((out - base).abs().max() / base.abs().max()):.3e}")
