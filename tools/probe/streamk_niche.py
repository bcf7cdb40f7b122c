"""tools/probe/streamk_niche.py -- the shapes the chooser still gives to STREAMK: several slabs of tiny rows that start off the 16-byte grid"""
import os, sys
import importlib.util
sys.argv = ["tiny_inner.py", "none"]
spec = importlib.util.spec_from_file_location("ti", os.path.join(os.path.dirname(os.path.abspath(__file__)), "tiny_inner.py"))
ti = importlib.util.module_from_spec(spec); spec.loader.exec_module(ti)
S = [{"TTV_B200_USE_STREAMK": "0"}, {}]
for dt, na in [("f32", [3, (1 << 20) + 1, 64]), ("f32", [5, (1 << 20) + 3, 48]), ("f32", [2, (1 << 20) + 1, 128]), ("f64", [3, (1 << 19) + 1, 64])]:
    ti.run(dt, na, 2, S)
