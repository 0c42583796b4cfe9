"""Drop-in for kameris/job_steps/__init__.py:9-15 restricted to the steps this package replaces
(the two hot-path steps, and `mds`, their first consumer: SURVEY.md 8f-4): merge `step_runners` into
kameris' own table."""
from . import backend
from . import mds

step_runners = {
    'distances': backend.run_backend_dists,
    'kmers': backend.run_backend_kmers,
    'mds': mds.run_mds_step,
}
