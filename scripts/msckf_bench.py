#!/usr/bin/env python3
"""Time the MSCKF (CTA-per-filter) fused step: bench.run_extras' MSCKF entry only."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
peak, _ = bench.measured_peaks()
dev = torch.device("cuda", 0)
# reuse the extras code path but only print the MSCKF entry
out = bench.run_extras(dev, peak)
print(json.dumps({k: v for k, v in out.items() if "msckf" in k or "rts" in k or "history" in k}, indent=1))
