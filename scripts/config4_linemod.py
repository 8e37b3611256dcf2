"""BASELINE config 4 (LINEMOD shape): dense SO(3) grid (>= 50k hypotheses per pair) + top-k refinement pass.

Synthetic task with a known answer: the target volume IS the source volume rotated by a hidden rotation, so
the verification score peaks at that rotation whatever the (random-init) head weights are.  Pass 1 scores a
deterministic super-Fibonacci grid and keeps the top-k; pass 2 scores m local perturbations of each
(`HypothesisVerifier.refine`).  Reports the geodesic error after each pass and the time per pair batch.
Prints one JSON line.   AHV_B (8), AHV_N (50000), AHV_K (32), AHV_M (64)."""
import importlib, json, os, statistics, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from modules._estimator_base import geodesic_deg
ahv = importlib.import_module("3dahv_b200")
dev = torch.device("cuda", 0)
B, N = int(os.environ.get("AHV_B", "8")), int(os.environ.get("AHV_N", "50000"))
k, m = int(os.environ.get("AHV_K", "32")), int(os.environ.get("AHV_M", "64"))
W1, W2, b2, vs, _, _ = bench.synthetic_inputs(torch, B, 16)
v = ahv.HypothesisVerifier(W1.to(dev), W2.to(dev), b2.to(dev))
vs = vs.to(dev)
R_gt = ahv.so3.sample_rotations(B, seed=123, device=dev)
vt = ahv.ops.rotate_volume(vs, R_gt)                       # per-pair volumes, one rotation each
grid = ahv.so3.grid_rotations(N, device=dev)
first = v.score(vs, vt, grid, k=1, return_scores=False)
err1 = geodesic_deg(first.R_best[:, 0], R_gt)
R2, s2, _, _ = v.refine(vs, vt, grid, k=k, m=m, max_angle_deg=6.0)
err2 = geodesic_deg(R2, R_gt)

def timed(fn, n=20):
    for _ in range(3):
        fn()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(n)]
    for a, b in ev:
        a.record(); fn(); b.record()
    torch.cuda.synchronize()
    return statistics.median(a.elapsed_time(b) for a, b in ev)

t1 = timed(lambda: v.score(vs, vt, grid, k=1, return_scores=False))
t2 = timed(lambda: v.refine(vs, vt, grid, k=k, m=m, max_angle_deg=6.0))
print(json.dumps({"what": "config 4: dense SO(3) grid + top-k refinement on a synthetic task with a known rotation",
                  "pairs": B, "grid_hypotheses": N, "top_k": k, "perturbations_per_candidate": m,
                  "geodesic_err_deg_pass1": {"mean": float(err1.mean()), "max": float(err1.max())},
                  "geodesic_err_deg_refined": {"mean": float(err2.mean()), "max": float(err2.max())},
                  "score_pass1_mean": float(first.topk_val.mean()), "score_refined_mean": float(s2.mean()),
                  "ms_pass1": t1, "ms_two_pass": t2,
                  "hyp_pairs_per_s_two_pass": B * (N + k * m) / (t2 * 1e-3)}))
