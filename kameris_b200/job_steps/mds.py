"""Drop-in for kameris/job_steps/mds.py: run_mds_step(options, exp_options) with the same option keys
(`dists_file`, `dimensions`, `output_file`; schema kameris/schemas/job_options.json:104-116) and the
same output (a JSON list of n points).  The distance triangle goes from the file to the device as it
is; centring and the eigen-solve run there (kameris_b200/mds.py); no CPU fallback."""
from __future__ import absolute_import, division, unicode_literals

import json
import logging
import time

from .. import formats


def run_mds_step(options, exp_options):
    import torch
    from .. import mds as M
    t0 = time.time()
    tri, n = formats.dist_reader.read_triangle(options['dists_file'])
    if tri.dtype.kind == 'u':
        tri = tri.astype('<u4').view('<i4')
    else:
        tri = tri.astype('<f4')
    dev = torch.device('cuda', torch.cuda.current_device())
    points = M.mds_from_triangle(torch.from_numpy(tri).to(dev), n, options['dimensions']).tolist()
    with open(options['output_file'], 'w') as outfile:
        json.dump(points, outfile)
    logging.getLogger('kameris').info('mds: %d points in %d dimensions in %.2f seconds', n,
                                      options['dimensions'], time.time() - t0)
