import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import oracle as O
from kameris_b200 import dist as D, _lib
from tests.helpers import synth_records, HIV_P
lib = _lib.load()
for n, L, k in [(4, 9000, 6), (2, 500, 3), (37, 3000, 5), (130, 9000, 6), (300, 2000, 4)]:
    recs = synth_records(n, L, seed=1, p=HIV_P)
    x = np.stack([O.cgr_count(r, k) for r in recs])
    xd = torch.from_numpy(x.view(np.int16)).cuda().view(torch.uint16)
    for cl in (1, 2):
        lib.kam_gram_set_cluster(cl)
        try:
            got = D.gram_tri(xd, ("cosine",))["cosine"].cpu().numpy()
            ref = O.cdist_triangle(x, "cosine")
            print(n, L, k, "cluster", cl, "max rel err", np.abs(got / ref - 1).max())
        except Exception as e:
            print(n, L, k, "cluster", cl, "ERROR", str(e)[:300]); sys.exit(1)
