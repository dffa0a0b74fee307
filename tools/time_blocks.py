"""Device-resident timing of the secondary blocks (CUDA events, inputs > L2). usage: time_blocks.py [xengine] [filter] [pfb] [fir] [mathconst]"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import bench
from gr_clenabled_b200 import blocks, capi

sp = torch.cuda.current_stream().cuda_stream
hbm, _ = bench.peaks()
res = bench.secondary_blocks(torch, blocks, capi, 0, sp, hbm)
for k, v in res.items():
    print(k, json.dumps(v))
