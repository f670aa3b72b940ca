# Small-problem latency probe (run on the GPU box from the repo root: PYTHONPATH=. python profiles/latency_probe.py)
import time, numpy as np, torch
from statdepth_b200._engine import get_engine
eng = get_engine()
rng = np.random.default_rng(0)
def bench(name, fn, reps=200):
    for _ in range(10): fn()
    t=time.perf_counter()
    for _ in range(reps): fn()
    dt=(time.perf_counter()-t)/reps
    print(f"{name}: {dt*1e6:.1f} us/call", eng.timings())
X = rng.standard_normal((100,200)).cumsum(0)
bench("cfg1 strict host", lambda: eng.band_depth_counts(X, None, 2, False))
bench("cfg1 relaxed host", lambda: eng.band_depth_counts(X, None, 2, True))
Xd = torch.from_numpy(X).cuda(); out = torch.zeros(200, dtype=torch.int64, device='cuda')
bench("cfg1 strict dev", lambda: eng.band_depth_counts_dev(Xd.data_ptr(), 100, 200, 200, out.data_ptr(), None, 200, 2, False))
bench("cfg1 relaxed dev", lambda: eng.band_depth_counts_dev(Xd.data_ptr(), 100, 200, 200, out.data_ptr(), None, 200, 2, True))
X2 = rng.standard_normal((128,100000)).cumsum(0)
X2d = torch.from_numpy(X2).cuda(); out2 = torch.zeros(100000, dtype=torch.int64, device='cuda')
bench("128 x 100k relaxed dev", lambda: eng.band_depth_counts_dev(X2d.data_ptr(), 128, 100000, 100000, out2.data_ptr(), None, 100000, 2, True), 100)
print(eng.phase_ns())
