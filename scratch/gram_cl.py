import os, sys, json
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from kameris_b200 import _lib, dist_bench
lib = _lib.load()
dev = torch.device("cuda", 0); torch.cuda.set_device(dev)
for cl in (1, 2):
    for f16 in (0, 1):
        lib.kam_gram_set_cluster(cl); lib.kam_gram_force_f16(f16)
        n, k = 8192, 9
        d = 4 ** k
        x = dist_bench.synth_counts(torch, dev, n, 9000, k, 20261003)
        pairs = n * (n - 1) // 2
        tri = [torch.empty(pairs, dtype=torch.float32, device=dev) for _ in range(3)]
        ws = torch.empty(lib.kam_gram_workspace_bytes(n, d), dtype=torch.uint8, device=dev)
        st = torch.cuda.current_stream(dev).cuda_stream
        def gram():
            _lib.check(lib.kam_gram_tri(x.data_ptr(), 1, n, d, x.stride(0), 0, n, 28, 1, tri[0].data_ptr(), tri[1].data_ptr(), tri[2].data_ptr(), ws.data_ptr(), ws.numel(), st))
        t = dist_bench._time(torch, gram, 3)
        print("cluster", cl, "f16" if f16 else "i8", "ms %.3f" % (t * 1e3), "pairs/s %.3e" % (pairs / t), "TF-equiv %.0f" % (2.0 * d * pairs / t / 1e12), "chk", float(tri[1][12345]))
        del x, tri, ws
lib.kam_gram_force_f16(0)
