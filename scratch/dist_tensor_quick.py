"""python -m kameris_b200.dist_bench, only the manhattan / info lines"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from kameris_b200 import dist_bench
dev = torch.device("cuda", 0); torch.cuda.set_device(dev)
r = dist_bench.run(torch, dev, 6545.0, 1653.1)
for k, v in r["rooflines"].items():
    print(k, json.dumps(v))
