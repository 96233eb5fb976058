#!/usr/bin/env python
"""Build an A/B variant of the library into variants/<name>.so (git-ignored): python tools/build_variant.py NAME -DFLAG=...
Load it with QUADSIM_B200_LIB=variants/NAME.so."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from quadsim_b200 import build

name, flags = sys.argv[1], sys.argv[2:]
os.makedirs("variants", exist_ok=True)
build.build(force=True, extra_flags=flags, out="variants/%s.so" % name)
print(name)
