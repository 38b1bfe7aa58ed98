#!/usr/bin/env python
"""Build an experimental variant of libb200cs.so for A/B kernel measurements.

    python tools/build_variant.py NAME [-DFLAG ...]   ->  build/variants/libb200cs_NAME.so

Run a tool or the bench against it with B200CS_LIB=build/variants/libb200cs_NAME.so."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from numbacs_b200 import _build  # noqa: E402

name, flags = sys.argv[1], sys.argv[2:]
vdir = os.path.join(ROOT, "build", "variants")
os.makedirs(vdir, exist_ok=True)
print(_build.build(force=True, extra_flags=flags, lib=os.path.join(vdir, f"libb200cs_{name}.so"),
                   objdir=os.path.join(vdir, "obj_" + name)))
