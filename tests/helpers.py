"""Shared synthetic-input generators (deterministic)."""
import numpy as np

ALPHA = np.frombuffer(b"ACGT", dtype=np.uint8)
IUPAC = np.frombuffer(b"NRYKMSWBDHVacgtn-", dtype=np.uint8)
HIV_P = np.array([0.36, 0.18, 0.24, 0.22])


def synth_records(n, length, seed, dirty=0.0, p=None, jitter=0):
    """n records of ~length bases; `dirty` = fraction of non-ACGT symbols (incl. one inside the
    first 8 characters so quirk Q1 is exercised)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    recs = []
    for _ in range(n):
        L = length if not jitter else int(rng.integers(max(0, length - jitter), length + jitter + 1))
        s = ALPHA[rng.choice(4, size=L, p=p)].copy()
        if dirty > 0 and L > 0:
            m = rng.random(L) < dirty
            s[m] = IUPAC[rng.integers(0, len(IUPAC), size=int(m.sum()))]
            if L > 3:
                s[rng.integers(0, min(L, 8))] = ord("N")
        recs.append(s.tobytes())
    return recs
