"""Scratch: builds tuning variants of libcvo_b200 into build/variants/ (git-ignored, travels with gpurun).
usage: build_variants.py name:DEF1,DEF2 ..."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cvo_rgbd_b200 import build
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
os.makedirs(os.path.join(root, "build", "variants"), exist_ok=True)
for spec in sys.argv[1:]:
    name, _, defs = spec.partition(":")
    out = os.path.join(root, "build", "variants", "libcvo_b200_%s.so" % name)
    build.build_library(force=True, verbose=True, out=out, defines=[d for d in defs.split(",") if d])
    print("built", out)
