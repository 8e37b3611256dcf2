"""One config-2-sized call for ncu: AHV_VOL=f32|bf16, AHV_B, AHV_N."""
import importlib, os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
ahv = importlib.import_module("3dahv_b200")
if os.environ.get("AHV_VARIANT_LIB"):   # scripts-only: profile an experiment build (the product reads no environment)
    ahv._lib.LIB_PATH = os.path.abspath(os.environ["AHV_VARIANT_LIB"])
dev = torch.device("cuda", 0)
B, N = int(os.environ.get("AHV_B", "32")), int(os.environ.get("AHV_N", "50000"))
W1, W2, b2, vs, vt, normals = bench.synthetic_inputs(torch, B, N)
v = ahv.HypothesisVerifier(W1.to(dev), W2.to(dev), b2.to(dev))
R = ahv.ops.rotations_from_normals(normals.to(dev))
vs = vs.to(dev).to(torch.bfloat16 if os.environ.get("AHV_VOL", "f32") == "bf16" else torch.float32)
vt = vt.to(dev)
for _ in range(4):
    r = v.score(vs, vt, R, k=1, return_scores=False)
torch.cuda.synchronize()
print("ok", r.topk_idx[:4, 0].tolist())
