"""bench.py's `model_step` section on its own (GSSD training step at batch B: the model's torch modules, gssd_forward, and
gssd_forward(backbone=True)): python tools/model_step_perf.py [B]"""
import json
import os
import sys
import types

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

import bench

B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
a = types.SimpleNamespace(gmax=5)
print(json.dumps(bench.time_model_step(a, torch, torch.device("cuda:0"), B), indent=1))
