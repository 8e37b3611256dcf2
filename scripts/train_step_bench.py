"""Training-step timing: scores + InfoNCE forward/backward, per-pair rotation sets.
AHV_B (12), AHV_N (9000), AHV_CHUNK (1024)."""
import importlib, os, sys, time, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
ahv = importlib.import_module("3dahv_b200")
dev = torch.device("cuda", 0)
B, N = int(os.environ.get("AHV_B", "12")), int(os.environ.get("AHV_N", "9000"))
chunk = int(os.environ.get("AHV_CHUNK", "1024"))
W1, W2, b2, vs, vt, _ = bench.synthetic_inputs(torch, B, 16)
W1, W2, b2 = (t.to(dev).requires_grad_(True) for t in (W1, W2, b2))
vs, vt = vs.to(dev).requires_grad_(True), vt.to(dev).requires_grad_(True)
gt = ahv.so3.sample_rotations(B, seed=1, device=dev)
Rs = torch.cat([gt[:, None], ahv.so3.sample_rotations(B * (N - 1), seed=2, device=dev).reshape(B, N - 1, 3, 3)], 1).contiguous()

def step():
    s = ahv.training.verification_scores(vs, vt, Rs, W1, W2, b2, chunk=chunk)
    loss = ahv.training.infonce_loss(s, Rs, gt, acc_thr_deg=15.0).mean()
    loss.backward()
    return loss

for _ in range(2):
    step()
torch.cuda.synchronize()
t0 = time.perf_counter()
n = 3
for _ in range(n):
    loss = step()
torch.cuda.synchronize()
dt = (time.perf_counter() - t0) / n
print(f"B={B} N={N} chunk={chunk}: {dt*1e3:.1f} ms per training step ({B*N/dt:.3e} hyp*pairs/s fwd+bwd), loss {float(loss):.4f}")
