#!/usr/bin/env python
"""Benchmark of the 3DAHV hypothesis-and-verification hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

Metric (BASELINE.json): rotation hypotheses scored per second x pairs.
Workload at any N: BASELINE config 2 per GPU — CO3D shape, B=32 pairs, 50 000
hypotheses (test_co3d.py:212), fp32 volumes, one shared rotation set.  With N>1
each rank scores its own 50 000-hypothesis shard of an N x 50 000 set for all 32
pairs (weak scaling) and the per-rank top-k lists are merged by one NCCL
all-gather + merge kernel inside the timed step.

A step = target features (forward_3d2d on the 32 target volumes) + fused
score kernel + top-k (+ all-gather/merge) + gather of the winning rotations.
`value` is timed with CUDA events on resident inputs, L2 flushed between
steps; `e2e` goes through the host-buffer C-ABI entry (ahv_predict_host) with
H2D/D2H inside the timed region.  One JSON line on stdout (rank 0).
"""
from __future__ import annotations

import argparse
import importlib
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

PAIRS, HYPS, TOPK = 32, 50000, 1
METRIC, UNIT = "rotation hypotheses scored/sec (x pairs)", "hyp*pairs/s"
GATHER_BYTES_PER_HYP = 512 * 8 * 16 * 4      # 262 144 B of smem gather per hypothesis (fp32)
FLOP_PER_HYP = 1_843_200
HBM_BYTES_PER_HYP = 40                       # 36 B rotation in + 4 B score out


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return {"hbm_gbs": d.get("hbm_gbs", 6650.0), "bf16_tflops": d.get("bf16_tflops", 1590.0),
                "bf16_tflops_sustained": d.get("bf16_tflops_sustained", 1400.0),
                "sm_max_mhz": d.get("sm_max_mhz", 1965.0), "source": "measured (MEASURED_PEAKS.json)"}
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "sm_max_mhz": 1965.0,
            "source": "fallback (B200_PROFILING.md)"}


def ncu_traffic():
    """dram__bytes_read.sum + dram__bytes_write.sum of the score kernel, per launch, from the committed
    `ncu --set full` capture of this same workload (profiles/); None if the summary is absent."""
    import csv
    import glob
    files = sorted(glob.glob(os.path.join(ROOT, "profiles", "r*_ncu_score_tc_summary.csv")))
    if not files:
        return None
    rd = wr = None
    for row in csv.reader(l for l in open(files[-1]) if not l.startswith("#")):
        if len(row) >= 4 and row[0] in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
            scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(row[1], 1.0)
            v = float(row[-1]) * scale
            rd, wr = (v, wr) if row[0].endswith("read.sum") else (rd, v)
    return None if rd is None or wr is None else rd + wr


class ClockSampler:
    """nvidia-smi clocks/throttle reasons sampled DURING the timed region."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.index)], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2]))
            except (ValueError, IndexError):
                continue
            for name, col in (("hw_slowdown", 4), ("hw_thermal_slowdown", 5), ("sw_thermal_slowdown", 6), ("sw_power_cap", 7)):
                if len(r) > col and r[col].lower().startswith("active"):
                    reasons.add(name)
        busy = [s for s in sm if s > 0]
        return {"sm_mhz": statistics.median(busy) if busy else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def synthetic_inputs(torch, pairs, hyps, seed=0):
    """Random-init head of the named architecture (Conv2d 384->32 no bias, ReLU,
    Conv2d 32->32; PyTorch default init, modules/modules.py:66-70), Gaussian
    volumes with the statistics forward_2d3d produces, restated random_rotations
    from torch's CPU generator."""
    torch.manual_seed(seed)
    head = torch.nn.Sequential(torch.nn.Conv2d(384, 32, 1, bias=False), torch.nn.ReLU(), torch.nn.Conv2d(32, 32, 1))
    W1 = head[0].weight.detach().reshape(32, 384).clone()
    W2 = head[2].weight.detach().reshape(32, 32).clone()
    b2 = head[2].bias.detach().clone()
    vol_src = torch.randn(pairs, 16, 8, 8, 8) * 1.12 - 0.18
    vol_tgt = torch.randn(pairs, 16, 8, 8, 8) * 1.12 - 0.18
    normals = torch.randn(hyps, 4)
    return W1, W2, b2, vol_src, vol_tgt, normals


def cpu_reference_rate(torch, sample_pairs, sample_hyps, steps=1, warmup=0):
    """The reference's PyTorch CPU path (restated in oracle/ahv_oracle.py with the
    same ATen calls) on all host threads, on a bounded sample of the workload."""
    from oracle import ahv_oracle as orc

    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    W1, W2, b2, vs, vt, normals = synthetic_inputs(torch, sample_pairs, sample_hyps)
    R = orc.rotations_from_normals_torch(normals)
    times = []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        s = orc.score_torch(vs, vt, R, W1, W2, b2)
        best = torch.max(s, dim=1)
        _ = R[best.indices]
        dt = time.perf_counter() - t0
        if i >= warmup:
            times.append(dt)
    per_step = sum(times) / len(times)
    return {"value": sample_pairs * sample_hyps / per_step, "unit": UNIT, "cores": cores, "kind": "port",
            "threads": torch.get_num_threads(),
            "sample": f"{sample_pairs} pairs x {sample_hyps} hypotheses of the config-2 workload, "
                      f"oracle.score_torch (F.affine_grid/F.grid_sample/conv1x1/normalize, the reference's ATen calls), "
                      f"{per_step:.2f} s per pass"}, per_step


def run_reference(args):
    import torch

    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    sp, sh = 2, 10000
    base, per_step = cpu_reference_rate(torch, sp, sh, steps=max(1, args.steps), warmup=min(args.warmup, 1))
    line = {"impl": "reference", "metric": METRIC, "value": base["value"], "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": min(args.warmup, 1), "ms_per_step": per_step * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            # same workload as the product arm; each timed step scores a bounded sample of it on the host cores
            "config": {"workload": (f"CO3D config 2 (BASELINE.json configs[1]): B={PAIRS} pairs x N={HYPS} hypotheses per GPU, "
                                    f"fp32 volumes, shared rotation set, top-1; N>1 = weak scaling over hypothesis shards "
                                    f"+ in-kernel NVLink exchange of the winners"), "pairs": PAIRS, "hypotheses_per_gpu": HYPS,
                       "math": "reference ATen calls on the host CPU", "sample_per_step": f"{sp} pairs x {sh} hypotheses"},
            "cpu_baseline": base,
            "e2e": {"value": base["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def run_ours(args):
    import torch
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the 3DAHV hot path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    ahv = importlib.import_module("3dahv_b200")
    lib = ahv._lib.lib()
    math = {"tc": ahv.MATH_TC, "fp32": ahv.MATH_FP32}[args.math]

    B, N, k = args.pairs, args.hyps, TOPK
    strong = args.config == 3
    if strong:
        # BASELINE config 3: 128 pairs, bf16 volumes, ONE 50 000-hypothesis set sharded over the ranks
        B = 128 if args.pairs == PAIRS else args.pairs
        total_hyps = args.hyps
        lo, hi = ahv.dist.shard_bounds(total_hyps, rank, world)
        N = hi - lo
        W1, W2, b2, vs_h, vt_h, normals_h = synthetic_inputs(torch, B, total_hyps)
        normals_h = normals_h[lo:hi].contiguous()
        vs_h = vs_h.bfloat16()
        shard_lo = lo
    else:
        W1, W2, b2, vs_h, vt_h, normals_h = synthetic_inputs(torch, B, N * world)
        # this rank's shard of the N*world rotation set (global index offset = rank*N)
        normals_h = normals_h[rank * N:(rank + 1) * N].contiguous()
        shard_lo = rank * N
    verifier = ahv.HypothesisVerifier(W1.to(dev), W2.to(dev), b2.to(dev), math=math)
    vs, vt = vs_h.to(dev), vt_h.to(dev)
    R = ahv.ops.rotations_from_normals(normals_h.to(dev))
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)    # > 126 MB L2
    launches = {"n": 0}
    peer = None
    if world > 1 and k == 1 and args.collective == "peer" and math != ahv.MATH_FP32:
        try:
            peer = ahv.dist.PeerExchange(B, dev)
        except RuntimeError as e:   # raised on every rank together: fall back to NCCL all-gather + merge
            if rank == 0:
                print(f"[bench] {e}", file=sys.stderr)

    def step():
        if world == 1:   # ahv_verify: target-feature prologue + fused score/arg-max/selection = 2 launches
            r = verifier.score(vs, vt, R, k=k, return_scores=False, idx_offset=0, gather=True)
            launches["n"] += 2
            return r.topk_val, r.topk_idx, r.R_best
        if peer is not None:   # fused: the scoring kernels exchange and merge their winners over NVLink peer memory
            val, idx, Rb = ahv.ops.verify_sharded(vs, vt, R, verifier.W1, verifier.W2, verifier.b2, shard_lo, rank, world,
                                                  peer.ptrs, math=math, workspace=verifier._workspace(B, N, 1, dev))
            launches["n"] += 2
            return val, idx, Rb
        r = verifier.score(vs, vt, R, k=k, return_scores=False, idx_offset=shard_lo, gather=False)
        launches["n"] += 2
        vals, idxs = ahv.dist.all_gather_topk(r.topk_val, r.topk_idx)
        val, idx = ahv.ops.topk_merge(vals, idxs)
        own = (idx >= shard_lo) & (idx < shard_lo + N)
        Rb = ahv.ops.gather_rotations(R, torch.where(own, idx, torch.full_like(idx, shard_lo)), shard_lo)
        launches["n"] += 2                       # merge + winner gather (NCCL kernels not counted)
        return val, idx, Rb

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        step()
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    launches["n"] = 0
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    barrier()
    for a, b_ in ev:
        flush.zero_()                            # evict L2 between timed steps (untimed)
        a.record()
        step()
        b_.record()
    barrier()
    total_ms = sum(a.elapsed_time(b_) for a, b_ in ev)
    n_launch = launches["n"]

    # dominant kernel alone: score kernel, no selection (k=0), CUDA events on its stream
    tgt = verifier.target_features(vt)
    ws = torch.empty(max(ahv.ops.workspace_bytes(B, N, 1), 16), dtype=torch.uint8, device=dev)
    kev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    for a, b_ in kev:
        flush.zero_()
        a.record()
        ahv.ops.score(vs, tgt, R, verifier.W1, verifier.W2, verifier.b2, k=0, math=math, return_scores=False, workspace=ws)
        b_.record()
    torch.cuda.synchronize()
    kernel_ms = sum(a.elapsed_time(b_) for a, b_ in kev) / args.steps

    # other arithmetic modes on the same workload (informational; the headline stays fp32 volumes + AHV_MATH_TC)
    other = {}
    if rank == 0 and not strong:
        def _rate(src, m):
            for _ in range(2):
                ahv.ops.score(src, tgt, R, verifier.W1, verifier.W2, verifier.b2, k=0, math=m, return_scores=False, workspace=ws)
            a2, b2_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a2.record()
            for _ in range(3):
                ahv.ops.score(src, tgt, R, verifier.W1, verifier.W2, verifier.b2, k=0, math=m, return_scores=False, workspace=ws)
            b2_.record()
            torch.cuda.synchronize()
            return B * N * 3 / (a2.elapsed_time(b2_) * 1e-3)
        try:   # informational: must never cost the run its JSON line
            other["bf16_volumes_tc"] = _rate(vs.bfloat16(), ahv.MATH_TC)
            other["fp32_volumes_tc_f16gather"] = _rate(vs, ahv.MATH_TC_F16GATHER)
        except Exception as e:  # noqa: BLE001
            other["error"] = repr(e)

    # measured shared-memory read peak (no such figure in MEASURED_PEAKS.json)
    import ctypes
    nbytes = ctypes.c_ulonglong(0)
    sms = torch.cuda.get_device_properties(dev).multi_processor_count
    out = torch.zeros(2 * sms, device=dev)
    st = torch.cuda.current_stream(dev).cuda_stream
    lib.ahv_diag_smem_read(out.data_ptr(), 2 * sms, 200, ctypes.byref(nbytes), st)
    torch.cuda.synchronize()
    a, b_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    lib.ahv_diag_smem_read(out.data_ptr(), 2 * sms, 4000, ctypes.byref(nbytes), st)
    b_.record()
    torch.cuda.synchronize()
    smem_peak_gbs = nbytes.value / (a.elapsed_time(b_) * 1e-3) / 1e9
    clocks = sampler.stop() if rank == 0 else None

    # p50 per-pair latency, B=1: CUDA-graph replay of the whole step, CUDA events
    latency = {}
    if rank == 0 and not strong:
        import statistics as _st
        for n_lat in (3000, 50000):
            try:   # informational as well
                gv = ahv.GraphedVerifier(verifier, 1, n_lat, k=1, device=dev)
                gv(vs[:1].float(), vt[:1], R[:n_lat])
                for _ in range(10):
                    gv()
                lev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(100)]
                for a, b_ in lev:
                    a.record(); gv(); b_.record()
                torch.cuda.synchronize()
                latency[f"N={n_lat}"] = {"p50_us": _st.median(a.elapsed_time(b_) for a, b_ in lev) * 1e3, "calls": 100}
            except Exception as e:  # noqa: BLE001
                latency[f"N={n_lat}"] = {"error": repr(e)}

    # e2e: host buffers through the C ABI (H2D + compute + D2H inside the timed region)
    vs_p, vt_p = vs_h.float().pin_memory(), vt_h.pin_memory()
    R_p = R.cpu().pin_memory()
    for _ in range(2):
        ahv.ops.predict_host(vs_p, vt_p, R_p, W1, W2, b2, k=k, math=math, device=dev)
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        _, val_h, idx_h, Rb_h = ahv.ops.predict_host(vs_p, vt_p, R_p, W1, W2, b2, k=k, math=math, device=dev)
    barrier()
    e2e_s = time.perf_counter() - t0
    h2d = vs_p.numel() * 4 * 2 + R_p.numel() * 4 + (32 * 384 + 32 * 32 + 32 + 8) * 4
    d2h = B * k * (4 + 8 + 36)

    t = torch.tensor([total_ms, e2e_s, kernel_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms, e2e_s, kernel_ms = t.tolist()
    if rank == 0:
        peaks = load_peaks()
        units_per_step = B * args.hyps if strong else B * N * world
        value = units_per_step * args.steps / (total_ms * 1e-3)
        hyp_per_s_kernel = B * N / (kernel_ms * 1e-3)              # one GPU's kernel
        gather_bytes = GATHER_BYTES_PER_HYP // 2 if strong else GATHER_BYTES_PER_HYP   # 16-bit staged volume: 256 B per voxel sample
        gather_gbs = hyp_per_s_kernel * gather_bytes / 1e9
        nominal_smem = sms * 128 * peaks["sm_max_mhz"] * 1e6 / 1e9
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": total_ms / args.steps, "higher_is_better": True,
            "scaling": "strong" if strong else "weak", "vs_baseline": None, "dtype": "bf16 volumes" if strong else "f32",
            "data": "synthetic",
            "dtype_detail": "fp32 volumes, trilinear gather / normalise / correlate in fp32; the two 1x1 convs use fp16 "
                            "operands (10-bit mantissa, TF32-equivalent) with fp32 accumulation on tcgen05; scores within "
                            "1.3e-4 relative of the reference's fp32 CPU path (gate 1e-3)",
            "config": {"workload": (f"Objaverse config 3 (BASELINE.json configs[2]): B={B} pairs, bf16 volumes, one set of "
                                    f"{args.hyps} hypotheses sharded over {world} GPU(s), winners exchanged {'by the scoring kernels over NVLink peer memory' if peer is not None else 'with an NCCL all-gather'}") if strong else
                                   (f"CO3D config 2 (BASELINE.json configs[1]): B={B} pairs x N={N} hypotheses per GPU, "
                                    f"fp32 volumes, shared rotation set, top-{k}; N>1 = weak scaling over hypothesis shards "
                                    f"+ in-kernel NVLink exchange of the winners"), "pairs": B, "hypotheses_per_gpu": N, "math": args.math,
                       "collective": ("none (1 GPU)" if world == 1 else "peer memory, fused into the scoring kernel" if peer is not None
                                      else "NCCL all-gather + merge kernel"),
                       "l2": "flushed between timed steps (256 MiB memset, untimed)"},
            "voxel_samples_per_s": value * 512,
            "clocks": clocks,
            "latency_p50_per_pair": latency,
            "other_modes_hyp_pairs_per_s": other,
            "e2e": {"value": units_per_step * args.steps / e2e_s, "unit": UNIT, "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": d2h, "api": "ahv_predict_host (C ABI, pinned host buffers)"},
            "gpu_launches": n_launch,
            "roofline": {
                "bound": "smem", "kernel": "score (fused rotate+head+correlate)", "achieved": gather_gbs,
                "peak": smem_peak_gbs, "unit": "GB/s", "frac": gather_gbs / smem_peak_gbs,
                "peak_source": "measured in this run (ahv_diag_smem_read, conflict-free LDS.128)",
                "nominal_peak": nominal_smem, "kernel_ms": kernel_ms, "traffic": ncu_traffic(), "traffic_note": "DRAM bytes per launch from the committed ncu --set full capture (profiles/); algorithmic HBM bytes per launch = %d" % (HBM_BYTES_PER_HYP * B * N),
                "algorithmic_bytes_per_unit": gather_bytes,
                "alt": {
                    "hbm": {"achieved": hyp_per_s_kernel * HBM_BYTES_PER_HYP / 1e9, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                            "frac": hyp_per_s_kernel * HBM_BYTES_PER_HYP / 1e9 / peaks["hbm_gbs"]},
                    "tensor": {"achieved": hyp_per_s_kernel * FLOP_PER_HYP / 1e12, "peak": peaks["bf16_tflops"], "unit": "TFLOP/s",
                               "frac": hyp_per_s_kernel * FLOP_PER_HYP / 1e12 / peaks["bf16_tflops"]},
                    "peaks": peaks["source"]},
            },
        }
        if world == 1 and not args.no_cpu:
            base, _ = cpu_reference_rate(torch, 2, 10000)
            line["cpu_baseline"] = base
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", choices=["ours", "reference"], default="ours")
    ap.add_argument("--collective", choices=["peer", "nccl"], default="peer",
                    help="N>1: winners exchanged by the scoring kernels through NVLink peer memory (default) or "
                         "NCCL all-gather + merge kernel")
    ap.add_argument("--math", choices=["tc", "fp32"], default=os.environ.get("AHV_BENCH_MATH", "tc"))
    ap.add_argument("--pairs", type=int, default=PAIRS)
    ap.add_argument("--hyps", type=int, default=HYPS)
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--config", type=int, choices=[2, 3], default=2,
                    help="2 (default): BASELINE configs[1], weak scaling; 3: configs[2], bf16 volumes, hypothesis set sharded (strong)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
