"""Scratch: time the distance-array kernels at the C1 shape and check them against the reference's C kernels."""
import sys, os, time
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mdanalysis_b200 import _cabi
from mdanalysis_b200.lib import distances as gd
from oracle import oracle as orc

dev = torch.device("cuda")
g = torch.Generator(device=dev); g.manual_seed(7)
n, F = 3341, 98
xyz = (torch.rand((F + 1, n, 3), generator=g, device=dev) * 50.0).contiguous()
ref = xyz[:1].expand(F, n, 3).contiguous(); conf = xyz[1:].contiguous()
outbuf = torch.empty((F, n, n), dtype=torch.float64, device=dev)
sout = torch.empty((F, n * (n - 1) // 2), dtype=torch.float64, device=dev)
lib = _cabi.load(); st = _cabi.current_stream(xyz)
boxes = {"none": (0, None), "ortho": (1, np.array([50, 50, 50], dtype=np.float32)),
         "triclinic": (2, gd._box_args(np.array([50, 50, 50, 70.53571, 109.48542, 70.518196], dtype=np.float32))[1])}
REPS = int(os.environ.get('DA_REPS', '5'))
def timed(fn, reps=5):
    reps = REPS
    fn(); torch.cuda.synchronize()
    if reps == 0: return float('nan')
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
for name, (mode, b) in boxes.items():
    ms = timed(lambda: _cabi.check(lib.mdagpu_distance_array(_cabi.tensor_ptr(ref), n, _cabi.tensor_ptr(conf), n, mode, _cabi.ptr(b), _cabi.tensor_ptr(outbuf), F, 1, st)))
    ms2 = timed(lambda: _cabi.check(lib.mdagpu_self_distance_array(_cabi.tensor_ptr(conf), n, mode, _cabi.ptr(b), _cabi.tensor_ptr(sout), F, 1, st)))
    print(f"{name:10s} distance_array {ms:7.3f} ms {F*n*n*8/ms/1e6:7.0f} GB/s   self {ms2:7.3f} ms {F*(n*(n-1)//2)*8/ms2/1e6:7.0f} GB/s", flush=True)
if REPS:
    print('fill (write-only) of the 8.75 GB output: %.0f GB/s' % (outbuf.numel() * 8 / timed(lambda: outbuf.fill_(1.0)) / 1e6))
else:
    sys.exit(0)
# correctness: odd shapes around the tile sizes, far-apart coordinates (general path), coincident atoms
orc.build()
rng = np.random.default_rng(3)
box6 = {"none": None, "ortho": np.array([50, 40, 30, 90, 90, 90], dtype=np.float32),
        "triclinic": np.array([50, 50, 50, 70.53571, 109.48542, 70.518196], dtype=np.float32)}
bad = 0
for (na, nb, spread) in [(3341, 3341, 50), (33, 1025, 50), (1, 1, 50), (70, 2049, 400), (257, 1023, 50), (40, 5000, 1e4)]:
    a = (rng.random((na, 3)) * spread - spread / 4).astype(np.float32)
    b = (rng.random((nb, 3)) * spread - spread / 4).astype(np.float32)
    b[: min(na, nb) // 2] = a[: min(na, nb) // 2]          # zero distances
    for name, b6 in box6.items():
        got = gd.distance_array(a, b, box=b6)
        want = orc.distance_array(a, b, b6, impl="reference")
        ok = np.array_equal(got, want)
        gs = gd.self_distance_array(b, box=b6)
        ws = orc.self_distance_array(b, b6, impl="reference")
        ok2 = np.array_equal(gs, ws)
        if not (ok and ok2):
            bad += 1
            print("MISMATCH", na, nb, spread, name, ok, ok2, np.abs(got - want).max() if not ok else 0)
print("correctness:", "OK" if bad == 0 else f"{bad} FAILED")
