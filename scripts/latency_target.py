"""B=1 latency breakdown target (run under ncu --metrics gpu__time_duration.sum)."""
import importlib, os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
ahv = importlib.import_module("3dahv_b200")
dev = torch.device("cuda", 0)
N = int(os.environ.get("AHV_N", "3000"))
W1, W2, b2, vs, vt, normals = bench.synthetic_inputs(torch, 1, N)
v = ahv.HypothesisVerifier(W1.to(dev), W2.to(dev), b2.to(dev))
R = ahv.ops.rotations_from_normals(normals.to(dev))
vs, vt = vs.to(dev), vt.to(dev)
for _ in range(6):
    v.score(vs, vt, R, k=1, return_scores=False)
torch.cuda.synchronize()
