#!/usr/bin/env python
"""tools/sweep.py -- BASELINE.json configs[4]: isolated loss-kernel sweep, K in {2,3,4} x C in {2,4,19}.

For every (K, C) it times, on one B200, with CUDA events:
  ours   dct_jsd_fwdbwd_f32 (weight * mean JSD + gradients w.r.t. logits, one launch) over N pixels,
         streamed in chunks of [Bc,C,1024,1024] so that logits + grads stay under --mem-gb;
  aten   the reference's op-by-op composition restated in stock PyTorch on the SAME GPU
         (softmax x K, simplex checks with their host syncs, JSD_2D, .mean(), autograd backward;
         generalframework/loss/loss.py:70-84,183-196, utils/utils.py:142-151) on one chunk --
         the meaningful speed-up denominator;  bench-only code.
Writes a markdown table + JSON lines to --out.  Not part of the product path.
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def aten_entropy(p):
    import torch
    assert torch.allclose(p.sum(1), torch.ones_like(p.sum(1)))      # simplex(): host sync
    return -1.0 * (p * (p + 1e-16).log()).sum(1)


def aten_jsd_step(logits):
    """The reference's composition for `JSD_2D([softmax(z)]).mean().backward()` in stock ATen ops."""
    import torch
    import torch.nn.functional as F
    probs = [F.softmax(z, 1) for z in logits]
    for p in probs:
        assert torch.allclose(p.sum(1), torch.ones_like(p.sum(1)))  # simplex(): host sync
    mean = probs[0]
    for p in probs[1:]:
        mean = mean + p
    mean = mean / len(probs)
    f_term = aten_entropy(mean)
    ent = 0
    for p in probs:
        ent = ent + aten_entropy(p)
    loss = (f_term - ent / len(probs)).mean()
    loss.backward()
    return loss


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--pixels", type=float, default=2 ** 26, help="pixels per (K,C) point (64Mi default; up to 2**30)")
    ap.add_argument("--mem-gb", type=float, default=24.0, help="cap for logits+grads of one chunk")
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "sweep"))
    ap.add_argument("--ks", default="2,3,4")
    ap.add_argument("--cs", default="2,4,19")
    ap.add_argument("--no-aten", action="store_true")
    args = ap.parse_args()
    import torch

    import dct_b200
    from dct_b200 import _lib, _runtime
    dev = torch.device("cuda:0")
    torch.cuda.set_device(dev)
    dct_b200.set_check_mode("deferred")
    h = _lib.lib()
    os.makedirs(args.out, exist_ok=True)
    H = W = 1024
    N = int(args.pixels)
    rows = []
    peak = 6650.0
    pk = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(pk):
        try:
            peak = float(json.load(open(pk))["hbm_gbs"])
        except Exception:
            pass
    for K in [int(x) for x in args.ks.split(",")]:
        for C in [int(x) for x in args.cs.split(",")]:
            per_img = 2 * K * C * 4 * H * W
            Bc = max(1, min(N // (H * W), int(args.mem_gb * 1e9 // per_img)))
            nchunks = max(1, N // (Bc * H * W))
            g = torch.Generator(device=dev).manual_seed(1234)
            z = [3 * torch.randn(Bc, C, H, W, device=dev, generator=g) for _ in range(K)]
            gr = [torch.empty_like(t) for t in z]
            total = torch.zeros(1, dtype=torch.float64, device=dev)
            st = _runtime.state(dev)
            n_chunk = Bc * H * W

            def ours():
                _lib.check(h.dct_jsd_fwdbwd_f32(_lib.ptr_array(z), K, C, Bc, H * W, _lib.IN_LOGITS, 1.0 / N, None,
                                                total.data_ptr(), _lib.ptr_array(gr), None, None, 0, None,
                                                st.workspace.data_ptr(), _runtime.stream_ptr(dev)), "jsd")
            for _ in range(3):
                ours()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(args.reps):
                for _c in range(nchunks):   # chunk (> L2 by far) re-used as the stream's next chunk
                    ours()
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / args.reps
            pix = nchunks * n_chunk
            gbs = pix * (2 * K * C * 4) / (ms * 1e-3) / 1e9
            row = {"K": K, "C": C, "pixels": pix, "chunk_images": Bc, "chunks": nchunks, "ours_ms": ms,
                   "ours_pix_per_s": pix / (ms * 1e-3), "ours_alg_GBps": gbs, "frac_of_peak": gbs / peak, "peak_GBps": peak}
            if not args.no_aten:
                # one chunk of at most 8 images (the composition keeps ~(13K+8) full-size temporaries alive)
                Ba = min(Bc, 8 if C < 19 else 2)
                za = [t[:Ba].clone().requires_grad_() for t in z]
                for _ in range(2):
                    aten_jsd_step(za)
                    for t in za:
                        t.grad = None
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                for _ in range(args.reps):
                    aten_jsd_step(za)
                    for t in za:
                        t.grad = None
                torch.cuda.synchronize()
                dt = (time.perf_counter() - t0) / args.reps
                row.update({"aten_images": Ba, "aten_pix_per_s": Ba * H * W / dt, "speedup_vs_aten_same_gpu":
                            row["ours_pix_per_s"] / (Ba * H * W / dt)})
                del za
            rows.append(row)
            print(json.dumps(row), flush=True)
            del z, gr
            torch.cuda.empty_cache()
    with open(os.path.join(args.out, "sweep.jsonl"), "w") as f:
        for r in rows:
            f.write(json.dumps(r) + "\n")
    with open(os.path.join(args.out, "sweep.md"), "w") as f:
        f.write("| K | C | pixels | ours Gpix/s | ours GB/s (algorithmic 2KC4) | frac of HBM peak | ATen-composition Gpix/s (same B200) | speed-up |\n")
        f.write("|---|---|---|---|---|---|---|---|\n")
        for r in rows:
            f.write(f"| {r['K']} | {r['C']} | {r['pixels']:.3g} | {r['ours_pix_per_s'] / 1e9:.2f} | {r['ours_alg_GBps']:.0f} | "
                    f"{r['frac_of_peak']:.3f} | {r.get('aten_pix_per_s', 0) / 1e9:.3f} | {r.get('speedup_vs_aten_same_gpu', 0):.1f}x |\n")
    dct_b200.raise_if_flagged()


if __name__ == "__main__":
    main()
