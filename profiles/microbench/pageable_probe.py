import sys, time, os
sys.path.insert(0, '/root/repo')
import numpy as np, torch
from mdanalysis_b200 import FrameBlock
from mdanalysis_b200.analysis.rdf import InterRDF
rng = np.random.default_rng(1)
pos = rng.random((1000, 33334, 3), dtype=np.float32) * np.float32(100.0)
box = np.array([100]*3+[90]*3, dtype=np.float32)
pin = torch.from_numpy(pos).pin_memory().numpy()
for name, arr in (("pageable", pos), ("pinned", pin)):
    blk = FrameBlock(arr, box)
    fn = lambda: InterRDF(blk, None, nbins=150, range=(0.0, 15.0)).run()
    fn(); fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(3): fn()
    torch.cuda.synchronize()
    print(name, (time.perf_counter() - t0) / 3 * 1e3, "ms", flush=True)
