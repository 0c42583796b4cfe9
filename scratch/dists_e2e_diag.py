"""why kam_dists_host is slower inside bench.py than alone: time it after each kind of earlier activity"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from kameris_b200 import _lib
from kameris_b200.dist_bench import synth_counts
import bench
dev = torch.device("cuda:0"); torch.cuda.set_device(dev)
lib = _lib.load()
ne, de = 16384, 4096
xe = synth_counts(torch, dev, ne, 9000, 6, 20261006).cpu().pin_memory()
pe = ne * (ne - 1) // 2
h_man = torch.empty(pe, dtype=torch.int32).pin_memory()
def t(label):
    torch.cuda.synchronize()
    _lib.check(lib.kam_dists_host(xe.data_ptr(), 1, ne, de, 1, h_man.data_ptr(), None, None, None, None))
    t0 = time.perf_counter()
    for _ in range(3):
        _lib.check(lib.kam_dists_host(xe.data_ptr(), 1, ne, de, 1, h_man.data_ptr(), None, None, None, None))
    print("%-40s %.1f ms" % (label, (time.perf_counter() - t0) / 3 * 1e3), flush=True)
t("alone")
big = torch.empty((2500, 4 ** 9), dtype=torch.float64).pin_memory()
t("after 5.2 GB pinned alloc")
chars, start = bench.synth_chars(2500, 9000, 1)
hc = torch.from_numpy(chars.copy()).pin_memory()
_lib.check(lib.kam_kmers_host(hc.data_ptr(), start.ctypes.data, 2500, 9, 16, _lib.KAM_OUT_FREQ_F64, big.data_ptr()))
t("after kam_kmers_host")
g = torch.empty((20000, 262144), dtype=torch.float32, device=dev); del g
t("after 21 GB torch alloc (cached)")
torch.cuda.empty_cache()
t("after empty_cache")
import concurrent.futures as cf
with cf.ThreadPoolExecutor(16) as pool:                      # busy host threads for a moment
    list(pool.map(lambda i: np.sort(np.random.rand(2_000_000)), range(32)))
t("after a burst of host threads")
a = np.random.rand(1500, 1500); (a @ a).sum()
t("after numpy BLAS")
