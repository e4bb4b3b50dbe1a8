"""GPU probe: eager-PyTorch baseline of bench.py on its own.  python tools/eager_probe.py"""
import os, sys, traceback
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
model = bench.build_model().cuda().eval()
for tf32 in (False, True):
    try:
        print("tf32" if tf32 else "fp32", bench.eager_gpu_rates(model.state_dict(), tf32))
    except Exception:
        traceback.print_exc()
        break
