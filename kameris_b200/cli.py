"""Command-line shims with the reference executables' argv (kameris/job_steps/backend.py:36-39,
:61-65), so `binary_path()` can point at kameris_b200/scripts/generation_{cgr,dists}:

    generation_cgr  cgr <input directory> <output filename> <k> <16|32>
    generation_dists <input file> <output file prefix> <manhat,info[,euclid,cosine,pearson]>

Exit code != 0 on any failure (the reference aborts on an uncaught exception)."""
import logging
import sys

from .job_steps import backend

USAGE_CGR = ("Usage: generation_cgr <cgr type: cgr/twocgr> <input directory> <output filename> <k> "
             "<bits per element: 16/32> <if twocgr: chars for: first,second>")
USAGE_DISTS = "Usage: generation_dists <input file> <output file prefix> [manhat,info,euclid,cosine,pearson]"


def main_cgr(argv):
    if len(argv) < 5:
        print(USAGE_CGR)
        return 1
    mode, directory, out, k, bits = argv[:5]
    if mode != 'cgr':
        print('only the "cgr" representation is built (kameris never invokes twocgr: backend.py:36)')
        return 1
    backend.run_backend_kmers({'fasta_output_dir': directory, 'output_file': out, 'k': int(k),
                               'bits_per_element': int(bits), 'mode': 'counts', 'disable_avx': False}, {})
    return 0


def main_dists(argv):
    if len(argv) < 3:
        print(USAGE_DISTS)
        return 1
    names = [m for m in backend.REFERENCE_METRICS + backend.GRAM_METRICS if m in argv[2]]   # substring match
    if not names:
        print(USAGE_DISTS)
        return 1
    backend.run_backend_dists({'input_file': argv[0], 'output_prefix': argv[1], 'distances': names,
                               'disable_avx': False}, {})
    return 0


def _maybe_init_process_group():
    """Under `python -m torch.distributed.run --nproc-per-node N -m kameris_b200.cli ...` (one process per
    GPU) the steps shard over the ranks: bind this rank to its GPU and join the NCCL group."""
    import os
    if int(os.environ.get('WORLD_SIZE', '1')) <= 1:
        return False
    import torch
    import torch.distributed as dist
    if dist.is_initialized():
        return False
    dev = torch.device('cuda', int(os.environ.get('LOCAL_RANK', '0')))
    torch.cuda.set_device(dev)
    dist.init_process_group('nccl', device_id=dev)
    return True


def main(argv=None):
    argv = list(sys.argv[1:] if argv is None else argv)
    logging.basicConfig(level=logging.INFO, format='%(message)s', stream=sys.stdout)
    if not argv or argv[0] not in ('generation_cgr', 'generation_dists'):
        print(USAGE_CGR)
        print(USAGE_DISTS)
        return 1
    joined = False
    try:
        joined = _maybe_init_process_group()
        return (main_cgr if argv[0] == 'generation_cgr' else main_dists)(argv[1:])
    except Exception as e:  # noqa: BLE001 -- the executables die with a message and a non-zero code
        print('error: %s' % (getattr(e, 'output', None) or e))     # CalledProcessError carries the cause in .output
        return 1
    finally:
        if joined:
            import torch.distributed as dist
            dist.destroy_process_group()


if __name__ == '__main__':
    sys.exit(main())
