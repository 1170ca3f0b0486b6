"""BASELINE configs[4] alone: 1 h of audio -> log-mel (80 / 128 bins) + dedup features, HBM roofline."""
import json, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
out = bench.frontend_leg(torch.device("cuda", 0), peaks)
for k, v in out.items():
    if isinstance(v, dict):
        print(f"{k:16s} {v['us_per_hour_of_audio']:8.1f} us/h  {v['achieved_GBs']:7.1f} GB/s  {100 * v['frac_of_hbm_peak']:5.1f}% of HBM peak")
