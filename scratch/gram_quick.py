import sys, numpy as np, torch
sys.path.insert(0, '.')
from oracle import oracle as O
from kameris_b200 import dist as D
from tests.helpers import synth_records, HIV_P
recs = synth_records(37, 3000, seed=1, p=HIV_P)
x = np.stack([O.cgr_count(r, 5) for r in recs])
xd = torch.from_numpy(x.view(np.int16)).cuda().view(torch.uint16)
got = D.gram_tri(xd, ("cosine",))["cosine"].cpu().numpy()
ref = O.cdist_triangle(x, "cosine")
print("max rel err", np.abs(got/ref-1).max(), got[:5], ref[:5])
