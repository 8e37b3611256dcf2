import importlib, os, sys, torch, statistics
sys.path.insert(0, "/root/repo")
import bench
ahv = importlib.import_module("3dahv_b200")
dev = torch.device("cuda", 0)
for B, N in ((32, 50000), (1, 50000), (1, 200000)):
    W1, W2, b2, vs, vt, normals = bench.synthetic_inputs(torch, B, N)
    v = ahv.HypothesisVerifier(W1.to(dev), W2.to(dev), b2.to(dev))
    R = ahv.ops.rotations_from_normals(normals.to(dev))
    vs, vt = vs.to(dev), vt.to(dev)
    for k in (1, 32):
        for _ in range(3): v.score(vs, vt, R, k=k, return_scores=False)
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(20)]
        for a, b in ev:
            a.record(); v.score(vs, vt, R, k=k, return_scores=False); b.record()
        torch.cuda.synchronize()
        print(f"B={B} N={N} k={k}: {statistics.median(a.elapsed_time(b) for a,b in ev):.3f} ms")
    if hasattr(v, "refine"):
        for _ in range(2): v.refine(vs, vt, R, k=32)
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(10)]
        for a, b in ev:
            a.record(); v.refine(vs, vt, R, k=32); b.record()
        torch.cuda.synchronize()
        print(f"B={B} N={N} refine(k=32): {statistics.median(a.elapsed_time(b) for a,b in ev):.3f} ms")
