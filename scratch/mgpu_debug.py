import os, sys, tempfile, subprocess, socket
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tests.helpers import synth_records
from tests.test_multi_gpu import _run
root = tempfile.mkdtemp()
fdir = os.path.join(root, "fasta"); os.mkdir(fdir)
recs = synth_records(61, 3000, seed=8, jitter=2500, dirty=0.002)
for i, r in enumerate(recs):
    open(os.path.join(fdir, "%05d.fasta" % i), "wb").write(b">s%d\n" % i + r + b"\n")
names = "manhat,info,euclid,cosine,pearson"
for tag, w in (("one", 1), ("two", 2)):
    _run(w, ["generation_cgr", "cgr", fdir, os.path.join(root, tag + ".mm-repr"), "6", "16"], root)
    _run(w, ["generation_dists", os.path.join(root, tag + ".mm-repr"), os.path.join(root, tag), names], root)
for f in ["%s.mm-repr"] + ["%s-" + m + ".mm-dist" for m in names.split(",")]:
    a = np.fromfile(os.path.join(root, f % "one"), dtype=np.uint8); b = np.fromfile(os.path.join(root, f % "two"), dtype=np.uint8)
    if len(a) != len(b):
        print(f, "LENGTH", len(a), len(b)); continue
    d = np.nonzero(a != b)[0]
    print(f, "bytes", len(a), "diff bytes", len(d), "first", d[:5], "last", d[-5:] if len(d) else "")
    if len(d) and "dist" in f:
        hdr = 16
        fa = np.frombuffer(a[hdr:].tobytes(), dtype=np.float32 if "manhat" not in f else np.uint32)
        fb = np.frombuffer(b[hdr:].tobytes(), dtype=np.float32 if "manhat" not in f else np.uint32)
        idx = np.nonzero(fa != fb)[0]
        print("  elements differing", len(idx), idx[:10], fa[idx[:5]], fb[idx[:5]])
