"""ctypes binding of libpamnet_sm100.so (C ABI: include/pamnet_b200.h).

The library is the product: there is no CPU or eager-PyTorch fallback.  If the shared object is
missing or a call fails, this module raises.
"""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libpamnet_sm100.so")

c_i32, c_i64, c_f32, c_vp, c_sz = C.c_int32, C.c_int64, C.c_float, C.c_void_p, C.c_size_t


class Config(C.Structure):   # pamnet_config_t
    _fields_ = [("dataset", c_i32), ("dim", c_i32), ("n_layer", c_i32), ("flow", c_i32), ("simple", c_i32),
                ("cutoff_l", c_f32), ("cutoff_g", c_f32)]


class Sizes(C.Structure):    # pamnet_sizes_t
    _fields_ = [("n_nodes", c_i64), ("n_graphs", c_i64), ("n_edges_g", c_i64), ("n_edges_l", c_i64),
                ("n_t2", c_i64), ("n_t1", c_i64)]


class SbfConsts(C.Structure):  # pamnet_sbf_consts_t
    _fields_ = [("zeros", C.c_double * 42), ("norm", C.c_double * 42)]


_PC, _PS, _PB = C.POINTER(Config), C.POINTER(Sizes), C.POINTER(SbfConsts)

# name -> (restype, argtypes); every symbol include/pamnet_b200.h declares
SIGNATURES = {
    "pamnet_abi_version": (c_i32, []),
    "pamnet_last_error": (C.c_char_p, []),
    "pamnet_param_count": (c_i32, [_PC]),
    "pamnet_param_total": (c_i64, [_PC]),
    "pamnet_param_offsets": (c_i32, [_PC, c_vp, c_vp]),
    "pamnet_radius_count": (c_i32, [c_vp, c_vp, c_i64, c_f32, c_i32, c_i32, c_vp, c_vp, c_vp, c_vp]),
    "pamnet_radius_fill": (c_i32, [c_vp, c_vp, c_i64, c_f32, c_i32, c_i32, c_vp, c_i64, c_vp, c_vp]),
    "pamnet_radius_grid_scratch_bytes": (c_sz, [c_i64, c_i64]),
    "pamnet_radius_grid_count": (c_i32, [c_vp, c_vp, c_i64, c_i64, c_f32, c_i32, c_i32, c_vp, c_sz, c_vp, c_vp, c_vp, c_vp]),
    "pamnet_radius_grid_fill": (c_i32, [c_vp, c_vp, c_i64, c_i64, c_f32, c_i32, c_i32, c_vp, c_vp, c_i64, c_vp, c_vp]),
    "pamnet_knn": (c_i32, [c_vp, c_vp, c_i64, c_i32, c_vp, c_vp, c_vp]),
    "pamnet_knn_edges_count": (c_i32, [c_vp, c_vp, c_i64, c_i32, c_f32, c_vp, c_vp, c_vp, c_vp]),
    "pamnet_knn_edges_fill": (c_i32, [c_vp, c_vp, c_i64, c_i32, c_f32, c_vp, c_i64, c_vp, c_vp]),
    "pamnet_edge_filter_count": (c_i32, [c_vp, c_i64, c_vp, c_f32, c_vp, c_vp, c_vp, c_vp]),
    "pamnet_edge_filter_fill": (c_i32, [c_vp, c_i64, c_vp, c_vp, c_i64, c_vp, c_vp]),
    "pamnet_triplet_scratch_bytes": (c_sz, [c_i64, c_i64]),
    "pamnet_triplet_count": (c_i32, [c_vp, c_i64, c_i64, c_vp, c_vp, c_vp]),
    "pamnet_triplet_fill": (c_i32, [c_vp, c_i64, c_i64, c_vp, c_i64, c_i64] + [c_vp] * 10 + [c_vp]),
    "pamnet_plan_bytes": (c_i32, [_PC, _PS, C.POINTER(c_sz), C.POINTER(c_sz)]),
    "pamnet_plan_count": (c_i32, [_PC, _PS, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp]),
    "pamnet_plan_fill": (c_i32, [_PC, _PS, c_vp, c_vp, c_vp, c_vp]),
    "pamnet_workspace_bytes": (c_sz, [_PC, _PS]),
    "pamnet_model_forward": (c_i32, [_PC, _PS, _PB, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_sz, c_i32, c_vp, c_vp, c_vp,
                                     c_vp]),
    "pamnet_model_backward": (c_i32, [_PC, _PS, _PB, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_sz, c_vp, c_vp, c_vp, c_vp,
                                      c_vp]),
    "pamnet_prepared_weights_bytes": (c_sz, [_PC]),
    "pamnet_prepare_weights": (c_i32, [_PC, c_vp, c_vp, c_vp]),
    "pamnet_comm_unique_id": (c_i32, [c_vp]),
    "pamnet_comm_init": (c_i32, [c_vp, c_i32, c_i32]),
    "pamnet_comm_enable": (c_i32, [c_i32]),
    "pamnet_comm_destroy": (c_i32, []),
    "pamnet_grad_buckets": (c_i32, [c_i32]),
    "pamnet_wait_grad_bucket": (c_i32, [c_i32, c_vp]),
    "pamnet_grad_bucket_range": (c_i32, [_PC, c_i32, C.POINTER(c_i64), C.POINTER(c_i64)]),
    "pamnet_loss": (c_i32, [c_vp, c_vp, c_i64, c_i32, c_vp, c_vp, c_vp]),
    "pamnet_collate": (c_i32, [c_vp, c_i64, c_vp, c_vp, c_vp, c_vp, c_vp, c_i64, c_vp, c_i64, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp]),
    "pamnet_scatter_add": (c_i32, [c_vp, c_vp, c_i64, c_i64, c_i64, c_vp, c_vp]),
    "pamnet_bessel_rbf": (c_i32, [c_vp, c_i64, c_vp, c_f32, c_vp, c_vp]),
    "pamnet_sbf_radial": (c_i32, [_PB, c_vp, c_i64, c_f32, c_vp, c_vp]),
    "pamnet_spherical_basis": (c_i32, [_PB, c_vp, c_vp, c_vp, c_i64, c_vp, c_vp]),
    "pamnet_linear": (c_i32, [c_vp, c_i64, c_i32, c_i32, c_vp, c_vp, c_i32, c_vp, c_vp]),
    "pamnet_gemm": (c_i32, [c_i32, c_vp, c_i32, c_vp, c_i32, c_vp, c_i32, c_i32, c_i32, c_i32, c_i32, c_vp, c_vp]),
    "pamnet_plan_build_scratch_bytes": (c_sz, [_PC, c_i64, c_i64, c_i64]),
    "pamnet_plan_build": (c_i32, [_PC, c_vp, c_vp, c_i64, c_i64, c_vp, c_i64, c_i32, c_vp, c_i64, c_vp, c_i64, c_vp, c_sz,
                                  c_vp, c_sz, c_vp, c_sz, _PS, c_vp, c_vp]),
    "pamnet_optimizer_step": (c_i32, [c_vp, c_vp, c_vp, c_vp, c_vp, c_i64, c_vp, c_i32, c_i64, c_f32, c_f32, c_f32, c_f32,
                                      c_f32, c_f32, c_f32, c_i32, c_vp, c_vp]),
    "pamnet_debug_ws_offset": (c_i64, [_PC, _PS, C.c_char_p, c_i32]),
    "pamnet_debug_plan_offset": (c_i64, [_PS, c_i32, C.POINTER(c_i32)]),
    "pamnet_debug_launch_count": (c_i64, []),
    "pamnet_debug_profile_begin": (None, []),
    "pamnet_debug_profile_end": (c_i32, [c_vp, c_vp, c_vp, c_vp]),
    "pamnet_debug_profile_timeline": (c_i32, [c_vp, c_vp, c_vp, c_vp, c_i32]),
    "pamnet_debug_tc_trace": (c_i32, [c_vp, c_i32]),
    "pamnet_debug_chain_trace": (c_i32, [c_vp, c_i32]),
}

_lib = None


class PamnetError(RuntimeError):
    pass


def load():
    """Load (once) and return the shared library with typed entry points."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise PamnetError(
            f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(nvcc, sm_100a). pamnet_b200 has no CPU / eager fallback.")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)      # AttributeError here = header / library mismatch
        fn.restype, fn.argtypes = res, args
    if lib.pamnet_abi_version() != 1:
        raise PamnetError("libpamnet_sm100.so ABI version mismatch")
    _lib = lib
    return lib


KERNEL_CLASSES = ("gemm_f32", "node_chain", "global_msg_fwd", "global_msg_bwd", "local_edge_fwd", "local_msg_fwd",
                  "local_msg_bwd", "local_trip_bwd", "node_grad_gather", "basis", "graph", "readout", "misc")


def launch_count():
    return int(load().pamnet_debug_launch_count())


def profile_begin():
    load().pamnet_debug_profile_begin()


def profile_end():
    """-> {class: (ms, launches, algorithmic_bytes, fp32-equivalent flops)} accumulated since profile_begin()."""
    n = len(KERNEL_CLASSES)
    ms, cnt, byt, flo = (C.c_double * n)(), (c_i64 * n)(), (C.c_double * n)(), (C.c_double * n)()
    got = load().pamnet_debug_profile_end(ms, cnt, byt, flo)
    assert got == n, "kernel class table out of sync with csrc/common.cuh"
    return {k: (ms[i], int(cnt[i]), byt[i], flo[i]) for i, k in enumerate(KERNEL_CLASSES)}


def profile_timeline(cap=4096):
    """-> list of (class name, stream tag, t0 ms, t1 ms) since profile_begin()."""
    cls, tag = (c_i32 * cap)(), (c_i32 * cap)()
    t0, t1 = (c_f32 * cap)(), (c_f32 * cap)()
    n = load().pamnet_debug_profile_timeline(cls, tag, t0, t1, cap)
    return [(KERNEL_CLASSES[cls[i]], int(tag[i]), float(t0[i]), float(t1[i])) for i in range(max(n, 0))]


def check(rc, what=""):
    if rc != 0:
        msg = load().pamnet_last_error().decode(errors="replace")
        raise PamnetError(f"{what or 'pamnet call'} failed (rc={rc}): {msg}")


def ptr(t):
    """Device (or host) address of a tensor, None -> NULL."""
    return None if t is None else t.data_ptr()
