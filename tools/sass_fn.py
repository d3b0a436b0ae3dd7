#!/usr/bin/env python
"""Developer tool: print the SASS of one kernel of libdct_b200.so (first function whose demangled name matches the regex).

    python tools/sass_fn.py 'tile_kernel<dct::JsdOp<3, true, 2, true>, 4, 2, 8, 4' [lib] > /tmp/k.sass
"""
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[2] if len(sys.argv) > 2 else os.path.join(ROOT, "deep-co-training-for-semi-supervised-image-segmentation_b200", "libdct_b200.so")
pat = re.compile(sys.argv[1])
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
blocks = re.split(r"(?m)^\s*Function : ", out)[1:]
names = [b.split("\n", 1)[0].strip() for b in blocks]
dem = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.split("\n")
for b, d in zip(blocks, dem):
    if pat.search(d):
        sys.stdout.write("Function : " + d + "\n" + b.split("\n", 1)[1])
        break
else:
    sys.exit("no function matches; candidates:\n" + "\n".join(d[:160] for d in dem if "tile_kernel" in d)[:4000])
