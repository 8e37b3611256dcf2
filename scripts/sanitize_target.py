"""Small end-to-end invocation of every kernel for compute-sanitizer (memcheck / racecheck)."""
import importlib, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
ahv = importlib.import_module("3dahv_b200")
dev = torch.device("cuda", 0)
g = dict(np.load(os.path.join(ROOT, "tests/golden/shared_n3000_b3.npz")))
w = dict(np.load(os.path.join(ROOT, "tests/golden/weights.npz")))
T = lambda a: torch.from_numpy(a).to(dev)
W = [T(w[k]) for k in ("W1", "W2", "b2")]
n = int(os.environ.get("AHV_SAN_N", "37"))
R = ahv.ops.rotations_from_normals(T(g["normals"][:n]))
Rs = ahv.so3.sample_rotations(n, seed=1, device=dev)
for math in (ahv.MATH_FP32, ahv.MATH_TC):
    v = ahv.HypothesisVerifier(*W, math=math)
    r = v.score(T(g["vol_src"]), T(g["vol_tgt"]), R, k=5)
    rp = v.score(T(g["vol_src"]), T(g["vol_tgt"]), torch.stack([R, Rs, R]).contiguous(), k=1)
    rb = v.score(T(g["vol_src"]).bfloat16(), T(g["vol_tgt"]), R, k=1)
    torch.cuda.synchronize()
    err = (r.scores.cpu().numpy() - g["scores"][:, :n]) / g["scores"][:, :n]
    print("math", math, "max rel err", np.abs(err).max())
rot = ahv.ops.rotate_volume(T(g["vol_src"])[0], R)
mv, mi = ahv.ops.topk_merge(torch.stack([r.topk_val, r.topk_val]), torch.stack([r.topk_idx, r.topk_idx + 100]))
torch.cuda.synchronize()
print("done", rot.shape, mi[0].tolist())
# training variant: exact backward, saved-activation forward, both saved-activation backward forms (tensor-core one
# with a rotation set and with shrunk matrices - the adjoint's exact path), fused InfoNCE
if os.environ.get("AHV_SAN_TRAIN", "1") != "0":
    B, Nt = 2, int(os.environ.get("AHV_SAN_NT", "19"))
    vs, vt = T(g["vol_src"][:B]), T(g["vol_tgt"][:B])
    tgt = ahv.ops.forward_3d2d(vt, *W)
    Rp = ahv.so3.sample_rotations(B * Nt, seed=3, device=dev).reshape(B, Nt, 3, 3).contiguous()
    gs = torch.randn(B, Nt, device=dev)
    g0 = ahv.ops.score_backward(vs, tgt, Rp, *W, gs)
    sc, h1, pinv = ahv.ops.score_train(vs, tgt, Rp, *W)
    g1 = ahv.ops.score_backward(vs, tgt, Rp, *W, gs, h1, pinv, ahv.MATH_FP32)
    g2 = ahv.ops.score_backward(vs, tgt, Rp, *W, gs, h1, pinv, ahv.MATH_TC)
    sc3, h13, pinv3 = ahv.ops.score_train(vs, tgt, (Rp * 0.35).contiguous(), *W)
    g3 = ahv.ops.score_backward(vs, tgt, (Rp * 0.35).contiguous(), *W, gs, h13, pinv3, ahv.MATH_TC)
    loss, dl = ahv.ops.infonce(sc, Rp, Rp[:, 0].contiguous(), 15.0)
    torch.cuda.synchronize()
    print("training", [f"{float((a - b).abs().max() / b.abs().max()):.1e}" for a, b in zip(g2, g1)], float(loss.sum()))
