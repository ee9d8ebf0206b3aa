"""ctypes binding of libsparsebev_b200.so (the C ABI declared in include/sparsebev_b200.h).

There is NO fallback: if the library is missing or a call fails, a RuntimeError is raised.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
SO_PATH = os.path.join(_HERE, 'csrc', 'libsparsebev_b200.so')

c_f32p = ctypes.POINTER(ctypes.c_float)
c_i32p = ctypes.POINTER(ctypes.c_int32)
c_i64p = ctypes.POINTER(ctypes.c_int64)
c_vp = ctypes.c_void_p
c_int = ctypes.c_int
c_vpp = ctypes.POINTER(c_vp)



class DenseLayer(ctypes.Structure):
    """struct sbev_dense_layer (include/sparsebev_b200.h)"""
    _fields_ = [('Wt', c_vp), ('ldw', c_int), ('K', c_int), ('N', c_int), ('bias', c_vp), ('ln_w', c_vp), ('ln_b', c_vp),
                ('residual', c_vp), ('flags', c_int), ('y', c_vp), ('ldy', c_int), ('W_hi', c_vp), ('W_lo', c_vp), ('Kpad', c_int), ('y_hi', c_vp), ('y_lo', c_vp), ('W_pack', c_vp)]


MAX_PEERS = 8


class PeerSegment(ctypes.Structure):
    """struct sbev_peer_segment (include/sparsebev_b200.h)"""
    _fields_ = [('src', c_vp), ('dst', c_vp * MAX_PEERS), ('bytes', ctypes.c_int64)]


# name -> argtypes; every function returns int (SBEV_OK = 0)
SIGNATURES = {
    'sbev_set_option': [ctypes.c_char_p, c_int],
    'sbev_dense_chain_fwd': [c_vp, c_int, c_int, c_int, ctypes.POINTER(DenseLayer), c_vp, c_vp, c_int, c_int, c_vp],
    'sbev_dense_chain_points_fwd': [c_vp, c_int, c_int, c_int, ctypes.POINTER(DenseLayer), c_vp, c_f32p, c_int, c_int, c_int, c_int, c_vp, c_vp, c_vp],
    'sbev_dense_chain_reduce_fwd': [c_vp, c_int, c_vp, c_vp, c_vp, c_vp, c_vp, c_int, c_int, ctypes.POINTER(DenseLayer), c_vp, c_vp, c_int, c_int, c_vp],
    'sbev_dense_chain_ws_debug': [c_vp],
    'sbev_dense_chain_ws_fwd': [c_vp, c_int, c_int, c_int, ctypes.POINTER(DenseLayer), c_vp, ctypes.c_longlong, c_vp, c_vp, c_int, c_int, c_vp],
    'sbev_dense_chain_ws_reduce_fwd': [c_vp, c_int, c_vp, c_vp, c_vp, c_vp, c_vp, c_int, c_int, ctypes.POINTER(DenseLayer), c_vp, ctypes.c_longlong,
                                       c_vp, c_vp, c_int, c_int, c_vp],
    'sbev_msmv_fwd': [c_vpp, c_i32p, c_int, c_vp, c_vp, c_int, c_int, c_int, c_int, c_int, c_vp, c_vp],
    'sbev_msmv_bwd': [c_vp, c_vpp, c_i32p, c_int, c_vp, c_vp, c_int, c_int, c_int, c_int, c_int,
                      c_vpp, c_vp, c_vp, c_vp],
    'sbev_msmv_bwd_det': [c_vp, c_vpp, c_i32p, c_int, c_vp, c_vp, c_int, c_int, c_int, c_int, c_int,
                          c_vpp, c_vp, c_vp, c_vp, ctypes.c_longlong, c_vp],
    'sbev_msmv_indices': [c_i32p, c_int, c_vp, c_int, c_int, c_int, c_int, c_vp, c_vp, c_vp, c_vp, c_vp],
    'sbev_sampling4d_fwd': [c_vpp, c_i32p, c_int, c_i64p, c_i64p, c_i64p, c_i64p,
                            c_vp, c_vp, c_vp, c_vp, c_vp,
                            c_int, c_int, c_int, c_int, c_int, c_int, c_int,
                            ctypes.c_float, ctypes.c_float, ctypes.c_float, c_vp, c_vp, c_vp],
    'sbev_sampling4d_window_fwd': [c_vpp, c_i32p, c_int, c_i64p, c_i64p, c_i64p, c_i64p,
                                   c_vp, c_vp, c_int, c_vp, c_vp, c_vp,
                                   c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int,
                                   ctypes.c_float, ctypes.c_float, ctypes.c_float, c_vp, c_vp, c_vp],
    'sbev_sampling4d_scatter_fwd': [c_vpp, c_i32p, c_int, c_i64p, c_i64p, c_i64p, c_i64p,
                                    c_vp, c_vp, c_int, c_vp, c_vp, c_vp,
                                    c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int,
                                    ctypes.c_float, ctypes.c_float, ctypes.c_float, c_vpp, c_int, c_vp, c_vp],
    'sbev_sampling4d_owner_fwd': [c_vpp, c_i32p, c_int, c_i64p, c_i64p, c_i64p, c_i64p,
                                  c_vp, c_vp, c_int, c_vp, c_vp, c_vp,
                                  c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int,
                                  ctypes.c_float, ctypes.c_float, ctypes.c_float, c_vpp, c_int, c_int, c_vp],
    'sbev_peer_exchange': [ctypes.POINTER(PeerSegment), c_int, c_int, c_int, c_vpp, c_vp, c_vp],
    'sbev_sasa_split_range_fwd': [c_vp, c_vp, c_int, c_vp, c_vp, c_int, c_vp, c_f32p, c_int, c_int, c_int, c_int, c_int, c_int, c_vp, c_vp],
    'sbev_dense_fwd': [c_vp, c_int, c_vp, c_int, c_vp, c_vp, c_vp, c_vp, c_int, c_int, c_int, c_int, c_vp, c_vp],
    'sbev_refine_bbox_fwd': [c_vp, c_vp, c_vp, c_int, c_int, c_int, c_int, c_vp, c_vp],
    'sbev_sample_points_fwd': [c_vp, c_vp, c_int, c_vp, c_int, c_f32p, c_int, c_int, c_int, c_vp, c_vp, c_vp],
    'sbev_sasa_fwd': [c_vp, c_int, c_vp, c_vp, c_int, c_vp, c_f32p, c_int, c_int, c_int, c_int, c_vp, c_vp],
    'sbev_sasa_split_fwd': [c_vp, c_vp, c_int, c_vp, c_vp, c_int, c_vp, c_f32p, c_int, c_int, c_int, c_int, c_vp, c_vp],
    'sbev_mix_fwd': [c_vp, c_vp, c_int, c_int, c_int, c_int, c_int, c_vp, c_vp, c_vp, c_vp],
    'sbev_split_bf16': [c_vp, ctypes.c_int64, c_vp, c_vp, c_vp],
    'sbev_gemm_bf16_tn': [c_vpp, c_vpp, c_int, c_vp, c_int, c_int, c_int, c_int, c_vp, c_vp],
    'sbev_gemm_bf16_tn_split': [c_vp, c_vp, c_vp, c_vp, c_vp, c_int, c_int, c_int, c_vp, c_vp, c_vp],
    'sbev_mix_presplit_fwd': [c_vp, c_vp, c_vp, c_int, c_int, c_int, c_int, c_int, c_vp, c_vp, c_vp, c_vp],
    'sbev_reduce_ln_fwd': [c_vp, c_int, c_vp, c_vp, c_vp, c_vp, c_int, c_int, c_vp, c_vp],
    'sbev_conv2d_nhwc_fwd': [c_vp, c_int, c_int, c_int, c_int, c_vp, c_int, c_int, c_int, c_int, c_int, c_vp, c_vp,
                             c_vp, c_int, c_int, c_int, c_vp, c_int, c_vp],
    'sbev_stem_conv_fwd': [c_vp, c_int, c_int, c_int, c_vp, c_vp, c_vp, c_vp, c_vp],
    'sbev_stem_conv_k_fwd': [c_vp, c_int, c_int, c_int, c_vp, c_int, c_vp, c_vp, c_vp, c_vp],
    'sbev_maxpool3x3s2_nhwc_fwd': [c_vp, c_int, c_int, c_int, c_int, c_vp, c_vp],
    'sbev_maxpool3x3s2_ex_nhwc_fwd': [c_vp, c_int, c_int, c_int, c_int, c_int, c_int, c_vp, c_vp],
    'sbev_ese_nhwc_fwd': [c_vp, c_int, c_int, c_int, c_int, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp],
    'sbev_subsample2_nhwc_fwd': [c_vp, c_int, c_int, c_int, c_int, c_vp, c_vp],
    'sbev_cast_bf16': [c_vp, ctypes.c_int64, c_vp, c_vp],
}

_lib = None


def exported_symbols():
    return sorted(list(SIGNATURES) + ['sbev_abi_version', 'sbev_last_error', 'sbev_get_option', 'sbev_msmv_bwd_det_workspace', 'sbev_ese_workspace_floats',
                                             'sbev_dense_chain_ws_blob_bytes'])


def load():
    """Load the library once.  Raises RuntimeError (never falls back) when it is absent."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(SO_PATH):
        raise RuntimeError(
            'sparsebev_b200: %s not found. Build it with `python -m sparsebev_b200.build` '
            '(or `python -c "import __graft_entry__ as g; g.build()"`). There is no CPU / eager fallback.' % SO_PATH)
    lib = ctypes.CDLL(SO_PATH)
    missing = [n for n in exported_symbols() if not hasattr(lib, n)]
    if missing:         # a stale build of an older source tree
        raise RuntimeError('sparsebev_b200: %s lacks %s -- rebuild it (`python sparsebev_b200/build.py -f`)' % (SO_PATH, ', '.join(missing)))
    lib.sbev_abi_version.restype = c_int
    lib.sbev_last_error.restype = ctypes.c_char_p
    lib.sbev_get_option.argtypes = [ctypes.c_char_p]
    lib.sbev_get_option.restype = c_int
    lib.sbev_msmv_bwd_det_workspace.argtypes = [c_i32p, c_int, c_int, c_int, c_int, c_int]
    lib.sbev_msmv_bwd_det_workspace.restype = ctypes.c_longlong
    lib.sbev_ese_workspace_floats.argtypes = [c_int, c_int, c_int, c_int]
    lib.sbev_ese_workspace_floats.restype = ctypes.c_longlong
    lib.sbev_dense_chain_ws_blob_bytes.argtypes = [c_int, ctypes.POINTER(DenseLayer)]
    lib.sbev_dense_chain_ws_blob_bytes.restype = ctypes.c_longlong
    for name, argtypes in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.argtypes = argtypes
        fn.restype = c_int
    _lib = lib
    return lib


launch_count = 0      # number of successful C-ABI launches (each enqueues exactly one of OUR kernels)


options_epoch = 0     # bumped by set_option: captured CUDA graphs bake the kernel variants in and must be re-captured


def set_option(name, value):
    global options_epoch
    rc = load().sbev_set_option(name.encode(), int(value))
    if rc != 0:
        raise RuntimeError('sbev_set_option(%s) failed' % name)
    options_epoch += 1


def get_option(name):
    v = load().sbev_get_option(name.encode())
    if v < 0:
        raise RuntimeError('sbev_get_option(%s): unknown option' % name)
    return v


def check(rc, what):
    global launch_count
    launch_count += 1
    if rc != 0:
        msg = load().sbev_last_error()
        raise RuntimeError('%s failed (code %d): %s' % (what, rc, msg.decode() if msg else '?'))


def ptr_array(ptrs):
    return (c_vp * len(ptrs))(*ptrs)


def i32_array(vals):
    return (ctypes.c_int32 * len(vals))(*vals)


def i64_array(vals):
    return (ctypes.c_int64 * len(vals))(*vals)


def f32_array(vals):
    return (ctypes.c_float * len(vals))(*vals)
