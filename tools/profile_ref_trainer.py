#!/usr/bin/env python
"""Developer tool (GPU box): where the reference trainer's iteration time goes, stock vs with the drop-ins installed.

Runs the unmodified ``CoTrainer._train_loop`` (oracle/ref_trainer.py harness, staged reference in baseline/_ref) for a few
iterations under cProfile and prints the top host-side entries by cumulative time; GPU work shows up under whichever call
synchronises (``.item()`` / ``.cpu()``).        python tools/profile_ref_trainer.py [c2|c1] [iters]
"""
import cProfile
import io
import os
import pstats
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle")):
    sys.path.insert(0, p)

import torch  # noqa: E402

import ref_trainer as rt  # noqa: E402

CFG = {"c1": dict(K=2, arch="enet", C=4, B=4, H=256, W=256), "c2": dict(K=3, arch="unet", C=4, B=32, H=256, W=256)}


def main():
    wl = sys.argv[1] if len(sys.argv) > 1 else "c2"
    iters = int(sys.argv[2]) if len(sys.argv) > 2 else 4
    dev = torch.device("cuda", 0)
    for drop, mode in ((False, None), (True, None), (True, "deferred")):
        kw = dict(iters=iters, train_jsd=True, train_adv=True, deterministic=False, warmup_iters=1, tf32=None, check_mode=mode,
                  **CFG[wl])
        rt.run_train_loop(dev, drop, **kw)      # warm-up run (cudnn autotune, module load)
        torch.cuda.synchronize()
        pr = cProfile.Profile()
        pr.enable()
        out = rt.run_train_loop(dev, drop, **kw)
        torch.cuda.synchronize()
        pr.disable()
        s = io.StringIO()
        pstats.Stats(pr, stream=s).sort_stats("cumulative").print_stats(45)
        print(f"==== {wl} {'with drop-ins' if drop else 'stock'} (checks: {mode or 'eager'}): {out['it_per_s']:.3f} it/s over "
              f"{out['iters']} iterations")
        lines = [ln.replace(ROOT + "/", "") for ln in s.getvalue().splitlines() if ln.strip()]
        lines = [ln for ln in lines if "site-packages/torch/nn/modules/module.py" not in ln]
        print("\n".join(ln[:200] for ln in lines[4:40]))
    # timing only (no profiler), 8 iterations each
    for drop, mode in ((False, None), (True, None), (True, "deferred")):
        kw = dict(iters=8, train_jsd=True, train_adv=True, deterministic=False, warmup_iters=1, tf32=None, check_mode=mode, **CFG[wl])
        out = rt.run_train_loop(dev, drop, **kw)
        print(f"timing {wl} {'with drop-ins' if drop else 'stock'} (checks: {mode or 'eager'}): {out['it_per_s']:.3f} it/s")


if __name__ == "__main__":
    main()
