"""Measure one build of the library (product or an experiment variant from `3dahv_b200/build.py -D... --out=...`):
score-kernel time at the config-2 size for fp32 and bf16 volumes, many-pairs corners, B=1 latency, and a parity
check against the golden scores - one JSON line.  The variant is selected HERE by pointing the loader at another
file before its first use (the product itself reads no environment variable):

    python scripts/variant_bench.py [path/to/lib_variant.so] [--quick]
"""
import importlib, json, os, statistics, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
ahv = importlib.import_module("3dahv_b200")
args = [a for a in sys.argv[1:] if not a.startswith("--")]
if args:
    ahv._lib.LIB_PATH = os.path.abspath(args[0])
quick = "--quick" in sys.argv
dev = torch.device("cuda", 0)
lib = ahv._lib.lib()
out = {"lib": os.path.relpath(ahv._lib.LIB_PATH, ROOT)}

g = dict(np.load(os.path.join(ROOT, "tests/golden/shared_n3000_b3.npz")))
w = dict(np.load(os.path.join(ROOT, "tests/golden/weights.npz")))
T = lambda a: torch.from_numpy(a).to(dev)
v = ahv.HypothesisVerifier(T(w["W1"]), T(w["W2"]), T(w["b2"]))
r = v.score(T(g["vol_src"]), T(g["vol_tgt"]), T(g["R"]), k=1, return_scores=True)
s = r.scores.cpu().numpy()
out["golden_relerr_f32"] = float(np.max(np.abs(s - g["scores"]) / np.abs(g["scores"])))
out["golden_top1_ok"] = bool(np.array_equal(r.topk_idx[:, 0].cpu().numpy(), s.argmax(1)))
rb = v.score(T(g["vol_src"]).bfloat16(), T(g["vol_tgt"]), T(g["R"]), k=1, return_scores=True)
out["golden_relerr_bf16"] = float(np.max(np.abs(rb.scores.cpu().numpy() - g["scores"]) / np.abs(g["scores"])))

flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def kernel_ms(B, N, dtype, reps=5):
    W1, W2, b2, vs, vt, normals = bench.synthetic_inputs(torch, B, N)
    ver = ahv.HypothesisVerifier(W1.to(dev), W2.to(dev), b2.to(dev))
    R = ahv.ops.rotations_from_normals(normals.to(dev))
    vs = vs.to(dev).to(dtype)
    tgt = ver.target_features(vt.to(dev))
    ws = torch.empty(ahv.ops.workspace_bytes(B, N, 1), dtype=torch.uint8, device=dev)
    run = lambda: ahv.ops.score(vs, tgt, R, ver.W1, ver.W2, ver.b2, k=0, return_scores=False, workspace=ws)
    for _ in range(3):
        run()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(reps)]
    for a, b in ev:
        flush.zero_(); a.record(); run(); b.record()
    torch.cuda.synchronize()
    return statistics.median(a.elapsed_time(b) for a, b in ev)


for name, dt in (("f32", torch.float32), ("bf16", torch.bfloat16)):
    ms = kernel_ms(32, 50000, dt)
    out[f"config2_{name}_ms"] = ms
    out[f"config2_{name}_rate"] = 32 * 50000 / (ms * 1e-3)
if not quick:
    for B, N in ((256, 1000), (4096, 64), (8192, 8)):
        for name, dt in (("f32", torch.float32), ("bf16", torch.bfloat16)):
            ms = kernel_ms(B, N, dt)
            out[f"B{B}_N{N}_{name}_rate"] = B * N / (ms * 1e-3)
    for n_lat in (3000, 50000):
        W1, W2, b2, vs, vt, normals = bench.synthetic_inputs(torch, 1, n_lat)
        ver = ahv.HypothesisVerifier(W1.to(dev), W2.to(dev), b2.to(dev))
        gv = ahv.GraphedVerifier(ver, 1, n_lat, k=1, device=dev)
        gv(vs.to(dev), vt.to(dev), ahv.ops.rotations_from_normals(normals.to(dev)))
        for _ in range(10):
            gv()
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(100)]
        for a, b in ev:
            a.record(); gv(); b.record()
        torch.cuda.synchronize()
        out[f"b1_n{n_lat}_p50_us"] = statistics.median(a.elapsed_time(b) for a, b in ev) * 1e3
print(json.dumps(out), flush=True)
