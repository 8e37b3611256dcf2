"""p50 per-pair latency (B=1) through CUDA-graph replay: AHV_NS="3000,50000", AHV_VOL=f32|bf16."""
import importlib, os, statistics, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
ahv = importlib.import_module("3dahv_b200")
dev = torch.device("cuda", 0)
ns = [int(x) for x in os.environ.get("AHV_NS", "3000,50000").split(",")]
W1, W2, b2, vs, vt, normals = bench.synthetic_inputs(torch, 1, max(ns))
v = ahv.HypothesisVerifier(W1.to(dev), W2.to(dev), b2.to(dev))
R = ahv.ops.rotations_from_normals(normals.to(dev))
vdt = torch.bfloat16 if os.environ.get("AHV_VOL", "f32") == "bf16" else torch.float32
vs, vt = vs.to(dev).to(vdt), vt.to(dev)
for n in ns:
    gv = ahv.GraphedVerifier(v, 1, n, k=1, device=dev, vol_dtype=vdt)
    gv(vs, vt, R[:n])
    for _ in range(20):
        gv()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(200)]
    for a, b in ev:
        a.record(); gv(); b.record()
    torch.cuda.synchronize()
    t = sorted(a.elapsed_time(b) * 1e3 for a, b in ev)
    print(f"N={n} p50={statistics.median(t):.1f}us p10={t[20]:.1f} p90={t[180]:.1f} idx={int(gv.out.topk_idx[0, 0])}")
