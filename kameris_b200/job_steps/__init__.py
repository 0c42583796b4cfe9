"""Drop-in for kameris/job_steps/__init__.py:9-15 restricted to the two steps this package
replaces: merge `step_runners` into kameris' own table."""
from . import backend

step_runners = {
    'distances': backend.run_backend_dists,
    'kmers': backend.run_backend_kmers,
}
