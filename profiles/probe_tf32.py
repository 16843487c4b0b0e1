"""tensor-pipe probe variants: FG_TF32_PROBE_N (128|256), FG_TF32_PROBE_COMMIT (a tcgen05.commit every n*4 MMAs), FG_TF32_PROBE_VARY"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import face_generator_b200 as fg
ctx = fg.Context(0, max_batch=8, channels=3)
print({k: os.environ.get(k) for k in ("FG_TF32_PROBE_N", "FG_TF32_PROBE_MODE")}, "%.1f TFLOP/s" % ctx.tf32_peak(40000))
