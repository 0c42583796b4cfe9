import torch, json, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from kameris_b200 import repr_bench, _lib
dev = torch.device("cuda", 0)
lib = _lib.load()
for mode in (2, 1, 3, 1):
    lib.kam_kmer_set_long_path(mode)
    r = repr_bench.run(dev, 6545.0)["long_sequences"]
    print(mode, json.dumps({k: r[k] for k in ("value", "ms")}), r["roofline"]["frac"], flush=True)
lib.kam_kmer_set_long_path(1)
