"""Fixed-cost attribution of one B=1 launch (diagnostics build: python 3dahv_b200/build.py -DAHV_TIMELINE --out=experiments/variants/lib_timeline.so, then AHV_VARIANT_LIB=that file)."""
import ctypes, importlib, os, sys, numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
ahv = importlib.import_module("3dahv_b200")
if os.environ.get("AHV_VARIANT_LIB"):   # scripts-only: point the loader at the diagnostics build
    ahv._lib.LIB_PATH = os.path.abspath(os.environ["AHV_VARIANT_LIB"])
    ahv._lib.SIGNATURES["ahv_diag_timeline"] = (ctypes.c_int, [ctypes.c_void_p])
dev = torch.device("cuda", 0)
N = int(os.environ.get("AHV_N", "3000"))
W1, W2, b2, vs, vt, normals = bench.synthetic_inputs(torch, 1, N)
v = ahv.HypothesisVerifier(W1.to(dev), W2.to(dev), b2.to(dev))
R = ahv.ops.rotations_from_normals(normals.to(dev))
vs, vt = vs.to(dev), vt.to(dev)
gv = ahv.GraphedVerifier(v, 1, N, k=1, device=dev)
gv(vs, vt, R)
for _ in range(10):
    gv()
torch.cuda.synchronize()
lib = ahv._lib.lib()
buf = (ctypes.c_ulonglong * (160 * 16))()
assert lib.ahv_diag_timeline(buf) == 0
t = np.frombuffer(buf, dtype=np.uint64).reshape(160, 16).astype(np.int64)[:148]
t0 = t[:, 0].min()
names = ["entry", "setup done", "l1max done", "vol staged", "tile0 gathered", "weights packed", "tile0 MMAs issued",
         "epi: prologue grid done", "epi: tile0 phase A", "epi: last phase B", "roles done", "exit"]
for i, n in enumerate(names):
    c = t[:, i] - t0
    print(f"{n:26s} min {c.min():7d} ns  median {int(np.median(c)):7d}  max {c.max():7d}")
e = t[:, 0] - t0
late = np.nonzero(e > 1000)[0]
print("late CTAs:", len(late), late.tolist()[:40], e[late].tolist()[:40])
