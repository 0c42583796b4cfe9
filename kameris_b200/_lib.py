"""ctypes binding of libkameris_b200.so (C ABI: include/kameris_b200.h).

There is NO CPU fallback: if the library is missing it is built with nvcc; if that fails, or a
call fails, the error is raised.
"""
import ctypes
import os

from . import build as _build

c_void_p, c_size_t, c_int, c_uint32, c_uint64 = (ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int,
                                                 ctypes.c_uint32, ctypes.c_uint64)

KAM_U8, KAM_U16, KAM_U32, KAM_U64, KAM_F32, KAM_F64 = range(6)
KAM_OUT_COUNTS, KAM_OUT_FREQ_F32, KAM_OUT_FREQ_F64 = 0, 1, 2
KAM_M_MANHAT, KAM_M_INFO, KAM_M_EUCLID, KAM_M_COSINE, KAM_M_PEARSON = 1, 2, 4, 8, 16
KAM_GRAM_U8, KAM_GRAM_F16 = 0, 1
KAM_SPARSE_DECLINED = 0xffffffff
KAM_PEER_HANDLE_BYTES = 64
KAM_PEER_NO_FLAG = 0xffffffff


class PeerPiece(ctypes.Structure):
    """kam_peer_piece (include/kameris_b200.h)"""
    _fields_ = [("d_dst", ctypes.c_void_p), ("d_src", ctypes.c_void_p), ("bytes", ctypes.c_uint64),
                ("chunk", ctypes.c_uint32), ("reserved", ctypes.c_uint32)]

# every symbol include/kameris_b200.h declares: (restype, argtypes)
SIGNATURES = {
    "kam_version": (c_int, []),
    "kam_last_error": (ctypes.c_char_p, []),
    "kam_device_info": (c_int, [ctypes.POINTER(c_int), ctypes.POINTER(c_size_t), ctypes.POINTER(c_size_t)]),
    "kam_planes_words": (c_size_t, [c_uint64]),
    "kam_pack_workspace_bytes": (c_size_t, [c_uint64, c_uint32]),
    "kam_pack_ascii": (c_int, [c_void_p, c_void_p, c_uint32, c_uint64, c_void_p, c_void_p, c_void_p, c_void_p,
                               c_size_t, c_void_p]),
    "kam_pack_set_fused": (None, [c_int]),
    "kam_kmer_workspace_bytes": (c_size_t, [c_uint32, c_int]),
    "kam_kmer_count": (c_int, [c_void_p, c_void_p, c_void_p, c_uint32, c_uint64, c_int, c_int, c_int, c_void_p,
                               c_uint64, c_void_p, c_size_t, c_void_p]),
    "kam_kmer_sweep": (c_int, [c_void_p, c_void_p, c_void_p, c_uint32, c_uint64, c_int, c_int, c_int, c_int, c_void_p,
                               c_void_p, c_size_t, c_void_p]),
    "kam_kmer_sweep_set_fused_top": (None, [c_int]),
    "kam_kmer_set_warp_path": (None, [c_int]),
    "kam_kmer_set_long_path": (None, [c_int]),
    "kam_kmer_set_long_one_pass": (None, [c_int]),
    "kam_kmer_set_word_scan": (None, [c_int]),
    "kam_cgr_pool": (c_int, [c_void_p, c_void_p, c_void_p, c_uint32, c_int, c_int, c_void_p, c_uint64, c_void_p,
                             c_uint64, c_void_p]),
    "kam_normalize": (c_int, [c_void_p, c_int, c_uint32, c_uint64, c_uint64, c_int, c_void_p, c_uint64, c_void_p]),
    "kam_kmers_host": (c_int, [c_void_p, c_void_p, c_uint32, c_int, c_int, c_int, c_void_p]),
    "kam_kmers_host_keep": (c_int, [c_void_p, c_void_p, c_uint32, c_int, c_int, c_int, c_void_p, c_void_p, c_uint64]),
    "kam_kmers_host_set_sparse": (None, [c_int]),
    "kam_host_set_threads": (None, [c_int]),
    "kam_host_threads": (c_int, []),
    "kam_sparse_expand": (c_int, [c_void_p, c_void_p, c_void_p, c_uint32, c_int, c_int, c_int, c_void_p]),
    "kam_kmer_count_sparse": (c_int, [c_void_p, c_void_p, c_void_p, c_uint32, c_int, c_void_p, c_void_p, c_void_p,
                                      c_void_p, c_size_t, c_void_p]),
    "kam_fasta_record0": (ctypes.c_int64, [c_void_p, c_size_t, c_void_p]),
    "kam_fasta_file_sizes": (c_int, [c_void_p, c_uint32, c_void_p, c_int]),
    "kam_fasta_read_files": (c_int, [c_void_p, c_uint32, c_void_p, c_void_p, c_void_p, c_int, c_void_p]),
    "kam_manhattan_workspace_bytes": (c_size_t, [c_uint32, c_uint64]),
    "kam_manhattan_workspace_bytes_levels": (c_size_t, [c_uint32, c_uint64, c_uint32]),
    "kam_dist_set_tensor": (None, [c_int]),
    "kam_manhattan_tri": (c_int, [c_void_p, c_int, c_uint32, c_uint64, c_uint64, c_uint32, c_uint32, c_void_p,
                                  c_void_p, c_size_t, c_void_p]),
    "kam_manhattan_rect": (c_int, [c_void_p, c_void_p, c_int, c_uint32, c_uint32, c_uint64, c_uint64, c_uint64,
                                   c_void_p, c_void_p, c_size_t, c_void_p]),
    "kam_manhattan_rect_f32": (c_int, [c_void_p, c_int, c_void_p, c_uint32, c_uint32, c_uint64, c_uint64, c_uint64,
                                       c_void_p, c_void_p, c_size_t, c_void_p]),
    "kam_info_workspace_bytes": (c_size_t, [c_uint32, c_uint64]),
    "kam_info_tri": (c_int, [c_void_p, c_int, c_uint32, c_uint64, c_uint64, c_uint32, c_uint32, c_void_p, c_void_p,
                             c_size_t, c_void_p]),
    "kam_gram_workspace_bytes": (c_size_t, [c_uint32, c_uint64]),
    "kam_gram_workspace_bytes_int": (c_size_t, [c_uint32, c_uint64]),
    "kam_gram_force_f16": (None, [c_int]),
    "kam_gram_set_float_recover": (None, [c_int]),
    "kam_gram_set_cluster": (None, [c_int]),
    "kam_gram_set_pair": (None, [c_int]),
    "kam_gram_set_limb": (None, [c_int]),
    "kam_gram_dpad": (c_uint64, [c_uint64]),
    "kam_gram_prepare": (c_int, [c_void_p, c_int, c_uint32, c_uint64, c_uint64, c_int, c_void_p, c_void_p, c_void_p,
                                 c_void_p, c_void_p]),
    "kam_gram_block_workspace_bytes": (c_size_t, [c_uint32, c_uint32]),
    "kam_gram_block": (c_int, [c_void_p, c_void_p, c_void_p, c_uint32, c_void_p, c_void_p, c_void_p, c_uint32, c_int,
                               c_uint64, c_int, c_uint32, c_uint32, ctypes.c_uint, c_int, c_void_p, c_void_p, c_void_p,
                               c_uint64, c_int, c_void_p, c_size_t, c_void_p]),
    "kam_gram_fused": (c_int, [c_void_p, c_void_p, c_void_p, c_uint32, c_void_p, c_void_p, c_void_p, c_uint32, c_int,
                               c_uint64, c_int, c_void_p, c_uint32, c_void_p, c_uint32, ctypes.c_uint, c_int, c_void_p,
                               c_void_p, c_void_p, c_uint64, c_int, c_void_p, c_size_t, c_void_p]),
    "kam_dist_colmax": (c_int, [c_void_p, c_int, c_uint32, c_uint64, c_uint64, c_void_p, c_void_p]),
    "kam_dist_layout": (c_int, [c_void_p, c_uint64, c_void_p, c_void_p, c_void_p, c_void_p]),
    "kam_dist_kpad": (c_uint64, [c_uint64]),
    "kam_dist_encode": (c_int, [c_void_p, c_int, c_uint32, c_uint64, c_uint64, c_void_p, c_uint64, c_void_p, c_void_p,
                                c_void_p, c_void_p, c_size_t, c_void_p]),
    "kam_thermo_fused": (c_int, [c_void_p, c_void_p, c_void_p, c_uint32, c_void_p, c_void_p, c_void_p, c_uint32, c_uint64,
                                 c_uint64, c_int, c_void_p, c_uint32, c_void_p, c_uint32, ctypes.c_uint, c_void_p,
                                 c_uint64, c_void_p, c_size_t, c_void_p]),
    "kam_gram_set_wait_probe": (None, [c_void_p]),
    "kam_gram_set_strip": (None, [c_int]),
    "kam_peer_alloc": (c_int, [c_size_t, ctypes.POINTER(c_void_p)]),
    "kam_peer_free": (c_int, [c_void_p]),
    "kam_peer_export": (c_int, [c_void_p, c_void_p]),
    "kam_peer_open": (c_int, [c_void_p, ctypes.POINTER(c_void_p)]),
    "kam_peer_close": (c_int, [c_void_p]),
    "kam_peer_copy": (c_int, [c_void_p, c_void_p, c_size_t, c_void_p]),
    "kam_peer_pull": (c_int, [c_void_p, c_uint32, c_void_p, c_void_p, c_uint32]),
    "kam_peer_signal": (c_int, [c_void_p, c_void_p]),
    "kam_gram_tri": (c_int, [c_void_p, c_int, c_uint32, c_uint64, c_uint64, c_uint32, c_uint32, ctypes.c_uint,
                             c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]),
    "kam_mds_workspace_bytes": (c_size_t, [c_uint32]),
    "kam_mds_center": (c_int, [c_void_p, c_int, c_uint32, c_void_p, c_void_p, c_size_t, c_void_p]),
    "kam_symm_workspace_bytes": (c_size_t, [c_uint32]),
    "kam_symm_set_tma": (None, [c_int]),
    "kam_symm_block": (c_int, [c_void_p, c_uint32, c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]),
    "kam_colstats_workspace_bytes": (c_size_t, [c_uint32, c_uint64]),
    "kam_col_stats": (c_int, [c_void_p, c_int, c_uint32, c_uint64, c_uint64, c_void_p, c_void_p, c_void_p, c_void_p,
                              c_size_t, c_void_p]),
    "kam_col_scale": (c_int, [c_void_p, c_int, c_uint32, c_uint64, c_uint64, c_void_p, c_void_p, c_uint64, c_void_p]),
    "kam_dists_host": (c_int, [c_void_p, c_int, c_uint32, c_uint64, ctypes.c_uint, c_void_p, c_void_p, c_void_p,
                               c_void_p, c_void_p]),
    "kam_dists_device": (c_int, [c_void_p, c_int, c_uint32, c_uint64, c_uint64, ctypes.c_uint, c_int, c_void_p, c_void_p,
                                 c_void_p, c_void_p, c_void_p]),
    "kam_host_release": (None, []),
}

_lib = None


class KamError(RuntimeError):
    pass


def load():
    """Load (building if needed) libkameris_b200.so.  Raises when it cannot be had."""
    global _lib
    if _lib is None:
        path = _build.LIB
        if not os.path.exists(path):
            _build.build()
        lib = ctypes.CDLL(path)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)      # AttributeError if the ABI is incomplete
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib


KAM_E_WORKSPACE = -3


def check(rc):
    if rc != 0:
        raise KamError("libkameris_b200 error %d: %s" % (rc, load().kam_last_error().decode(errors="replace")))
