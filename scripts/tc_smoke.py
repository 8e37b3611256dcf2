"""Quick on-GPU check of the tensor-core path against the fp32 CUDA-core path and the goldens."""
import importlib, os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
ahv = importlib.import_module("3dahv_b200")
dev = torch.device("cuda", 0)
g = dict(np.load(os.path.join(ROOT, "tests/golden/shared_n3000_b3.npz")))
w = dict(np.load(os.path.join(ROOT, "tests/golden/weights.npz")))
T = lambda a: torch.from_numpy(a).to(dev)
W = [T(w[k]) for k in ("W1", "W2", "b2")]
for n in (2, 7, 64, 3000):
    out = {}
    for name, math in (("fp32", ahv.MATH_FP32), ("tc", ahv.MATH_TC)):
        v = ahv.HypothesisVerifier(*W, math=math)
        r = v.score(T(g["vol_src"]), T(g["vol_tgt"]), T(g["R"][:n]), k=1)
        torch.cuda.synchronize()
        out[name] = r.scores.cpu().numpy()
    ref = g["scores"][:, :n]
    for name in out:
        err = np.abs(out[name] - ref) / np.abs(ref)
        print(f"N={n} {name}: max rel err {err.max():.3e} mean {err.mean():.3e}", flush=True)
    if n <= 7:
        print(" tc  ", out["tc"][0, :7]); print(" ref ", ref[0, :7])
