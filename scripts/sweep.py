"""BASELINE config 5: hypotheses x pairs sweep on one GPU (+ B=1 latency), JSON lines to stdout.

    python scripts/sweep.py [--quick] > gpurun_out/sweep.jsonl
"""
import argparse, importlib, json, os, statistics, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402

ap = argparse.ArgumentParser(); ap.add_argument("--quick", action="store_true"); ap.add_argument("--math", default="tc")
args = ap.parse_args()
ahv = importlib.import_module("3dahv_b200")
dev = torch.device("cuda", 0)
math = {"tc": ahv.MATH_TC, "fp32": ahv.MATH_FP32, "tc_f16gather": ahv.MATH_TC_F16GATHER}[args.math]
Ns = [1000, 10000, 100000] if args.quick else [1000, 10000, 100000, 1000000]
Bs = [1, 16, 256] if args.quick else [1, 4, 16, 64, 256]
W1, W2, b2, vs_all, vt_all, normals = bench.synthetic_inputs(torch, max(Bs), max(Ns))
v = ahv.HypothesisVerifier(W1.to(dev), W2.to(dev), b2.to(dev), math=math)
R_all = ahv.ops.rotations_from_normals(normals.to(dev))
for B in Bs:
    for N in Ns:
        if B * N > 64_000_000:
            continue
        vs, vt, R = vs_all[:B].to(dev), vt_all[:B].to(dev), R_all[:N].contiguous()
        for dt, name in ((torch.float32, "f32"), (torch.bfloat16, "bf16")):
            src = vs.to(dt)
            for _ in range(3):
                v.score(src, vt, R, k=1, return_scores=False)
            reps = max(3, min(50, int(2e8 // (B * N))))
            ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(reps)]
            for a, b in ev:
                a.record(); v.score(src, vt, R, k=1, return_scores=False); b.record()
            torch.cuda.synchronize()
            ms = statistics.median(a.elapsed_time(b) for a, b in ev)
            print(json.dumps({"pairs": B, "hyps": N, "vol": name, "math": args.math, "ms_p50": ms,
                              "hyp_pairs_per_s": B * N / (ms * 1e-3), "voxel_samples_per_s": B * N * 512 / (ms * 1e-3)}), flush=True)
# B=1 latency with a CUDA graph (includes target features, scoring, selection, winner gather)
for N in (3000, 50000):
    gv = ahv.GraphedVerifier(v, 1, N, k=1, device=dev)
    gv(vs_all[:1].to(dev), vt_all[:1].to(dev), R_all[:N])
    for _ in range(20):
        gv()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(200)]
    for a, b in ev:
        a.record(); gv(); b.record()
    torch.cuda.synchronize()
    t = sorted(a.elapsed_time(b) for a, b in ev)
    print(json.dumps({"latency_B1": True, "hyps": N, "p50_us": t[100] * 1e3, "p90_us": t[180] * 1e3, "min_us": t[0] * 1e3}), flush=True)
