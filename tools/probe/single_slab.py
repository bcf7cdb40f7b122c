"""tools/probe/single_slab.py -- COLF on ONE (or a few) long slab(s) of tiny rows: how many partitions (the reduce pass sums them)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))


sys.argv = ["tiny_inner.py", "none"]
import importlib.util
spec = importlib.util.spec_from_file_location("ti", os.path.join(os.path.dirname(os.path.abspath(__file__)), "tiny_inner.py"))
ti = importlib.util.module_from_spec(spec); spec.loader.exec_module(ti)
S = [{"TTV_B200_USE_COLF": "0", "TTV_B200_USE_STREAMK": "0"}, {"TTV_B200_USE_COLF": "0", "TTV_B200_USE_STREAMK": "1"}, {}, {"TTV_B200_COLF_ITEMS_PER_WARP": "1"},
     {"TTV_B200_COLF_ITEMS_PER_WARP": "2"}, {"TTV_B200_KSPLIT": "1776"}, {"TTV_B200_KSPLIT": "3552"}]
for dt, na in [("f32", [2, (1 << 26) + 1]), ("f32", [2, 1 << 26]), ("f32", [3, 1 << 26]), ("f32", [2, 1 << 28]), ("f32", [2, 1 << 24, 4]), ("f32", [3, 1 << 24, 4])]:
    ti.run(dt, na, 2, S)
