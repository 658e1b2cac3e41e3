"""ctypes binding of libvae2_b200.so (the C ABI declared in include/vae2_b200.h).

There is NO fallback: if the shared library is missing or an entry point fails, the
caller gets a RuntimeError (same error behaviour as the reference's native-op precedent,
lib/models/sync_bn/inplace_abn/functions.py:25-28 ``_check`` -> RuntimeError).
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(os.path.dirname(_HERE), "libvae2_b200.so")

vp, i32, i64, f32 = C.c_void_p, C.c_int, C.c_int64, C.c_float
ip = C.POINTER(C.c_int)


class ConvGeom(C.Structure):
    _fields_ = [(n, C.c_int32) for n in
                ("B", "H", "W", "Cin_p", "ldx", "Ho", "Wo", "Cout_p", "ldy", "k", "stride", "pad")]


class PackDesc(C.Structure):
    _fields_ = [("w", vp), ("wp", vp), ("wpT", vp), ("wq", vp), ("wqT", vp), ("cin_map", vp),
                ("Cout", C.c_int32), ("Cin", C.c_int32), ("k", C.c_int32), ("Cin_p", C.c_int32),
                ("Cout_p", C.c_int32), ("reserved", C.c_int32)]


class Tf32PackDesc(C.Structure):
    _fields_ = [("w", vp), ("cin_map", vp), ("fwd", vp), ("bwd", vp),
                ("Cout", C.c_int32), ("Cin", C.c_int32), ("k", C.c_int32), ("Nf", C.c_int32), ("Kf", C.c_int32),
                ("NfT", C.c_int32), ("KfT", C.c_int32), ("reserved", C.c_int32)]


class FuseSrc(C.Structure):
    _fields_ = [("ptr", vp), ("H", C.c_int32), ("W", C.c_int32), ("ld", C.c_int32)]


class FuseDst(C.Structure):
    _fields_ = [("ptr", vp), ("ld", C.c_int32), ("accumulate", C.c_int32)]


class ElboSeg(C.Structure):
    _fields_ = [("kind", C.c_int32), ("slot", C.c_int32), ("a", vp), ("b", vp), ("out", vp),
                ("target", f32), ("scale", f32), ("Z", C.c_int32), ("HW", C.c_int32), ("n", i64),
                ("prior", C.c_int32), ("reserved", C.c_int32)]


class ElboBwdSeg(C.Structure):
    _fields_ = [("kind", C.c_int32), ("reserved0", C.c_int32), ("a", vp), ("b", vp), ("gz", vp), ("grad", vp),
                ("gout", vp), ("target", f32), ("scale", f32), ("Z", C.c_int32), ("HW", C.c_int32), ("n", i64),
                ("accumulate", C.c_int32), ("prior", C.c_int32)]


# name -> argtypes (restype is int status unless listed in _SPECIAL)
_PROTOS = {
    "vae2_nchw_to_act": [vp, vp, i32, i32, i32, i32, i32, i32, i32, i32, i32, vp],
    "vae2_act_to_nchw": [vp, vp, i32, i32, i32, i32, i32, i32, i32, i32, i32, vp],
    "vae2_slice_copy": [vp, vp, i32, i64, i32, i32, i32, i32, vp],
    "vae2_code_broadcast": [vp, vp, i32, i32, i32, i32, i32, i32, i32, vp],
    "vae2_spatial_sum": [vp, vp, i32, i32, i32, i32, i32, i32, i32, f32, i32, vp],
    "vae2_spatial_bcast": [vp, vp, i32, i32, i32, i32, i32, i32, i32, i32, f32, i32, vp],
    "vae2_pack_weights": [vp, i32, vp],
    "vae2_unpack_wgrad": [vp, i32, i32, vp],
    "vae2_pack_weights_tf32": [vp, i32, vp],
    "vae2_conv2d_fwd": [vp, vp, vp, vp, i32, C.POINTER(ConvGeom), i32, vp],
    "vae2_conv2d_dgrad": [vp, vp, vp, i32, C.POINTER(ConvGeom), i32, i32, vp],
    "vae2_conv2d_wgrad": [vp, vp, vp, i32, C.POINTER(ConvGeom), i32, vp],
    "vae2_bias_grad": [vp, vp, i32, i64, i32, i32, i32, vp],
    "vae2_conv2d_wgrad_tc": [vp, vp, vp, vp, C.POINTER(ConvGeom), vp],
    "vae2_conv2d_tc_supported": [C.POINTER(ConvGeom)],
    "vae2_conv2d_wgrad_f32x2": [vp, vp, vp, vp, C.POINTER(ConvGeom), vp],
    "vae2_bn_stats": [vp, vp, ip, i32, i64, i32, i32, vp],
    "vae2_bn_merge": [vp, i32, i32, vp, vp],
    "vae2_bn_finalize": [vp, i32, i32, i32, vp, vp, vp, vp, vp, f32, f32, vp, vp, vp, vp, vp],
    "vae2_bn_finalize_strided": [vp, i32, i64, i32, i32, vp, vp, vp, vp, vp, f32, f32, vp, vp, vp, vp, vp],
    "vae2_bn_eval_coeffs": [i32, i32, vp, vp, vp, vp, f32, vp, vp, vp],
    "vae2_bn_apply": [vp, vp, vp, i32, i64, i32, i32, i32, i32, vp, vp, i32, vp],
    "vae2_bn_bwd_reduce": [vp, vp, vp, vp, ip, i32, i64, i32, i32, i32, i32, vp, vp, i32, vp],
    "vae2_bn_bwd_finalize": [vp, i32, i32, i32, vp, vp],
    "vae2_bn_bwd_coeffs": [vp, i32, i32, f32, vp, vp, i32, vp, vp, vp, vp],
    "vae2_bn_bwd_elemt": [vp, vp, vp, vp, vp, i32, i64, i32, i32, i32, i32, i32, i32, vp, vp, vp, vp, vp,
                          i32, i32, i32, vp],
    "vae2_debug_bn_phase_times": [vp],
    "vae2_bn_fwd_fused": [vp, vp, vp, vp, i32, i64, i32, i32, i32, i32, i32, vp, vp, vp, vp, vp, f32, f32, vp, vp, vp,
                          vp, i32, vp],
    "vae2_bn_bwd_fused": [vp, vp, vp, vp, vp, vp, i32, i64, i32, i32, i32, i32, i32, i32, i32, vp, vp, vp, vp, vp, vp,
                          i32, vp, vp, i32, i32, i32, vp],
    "vae2_bn_fwd_fused_groups": [vp, vp, vp, vp, i32, i64, i32, i32, i32, i32, i32, vp, vp, vp, vp, vp, f32, f32, vp, vp, vp,
                                 vp, i32, i32, i32, vp],
    "vae2_bn_bwd_fused_groups": [vp, vp, vp, vp, vp, vp, i32, i64, i32, i32, i32, i32, i32, i32, i32, vp, vp, vp, vp, vp, vp,
                                 i32, vp, vp, i32, i32, i32, i32, i32, vp],
    "vae2_bn_sync_fwd_stats": [vp, vp, i32, i64, i32, i32, i32, i32, vp, vp],
    "vae2_bn_sync_fwd_apply": [vp, vp, vp, i32, i64, i32, i32, i32, i32, i32, vp, vp, vp, vp, vp, f32, f32, vp, vp, vp, vp,
                               i32, i32, i32, vp, i32, i64, vp],
    "vae2_bn_sync_bwd": [i32, vp, vp, vp, vp, vp, vp, i32, i64, i32, i32, i32, i32, i32, i32, i32, vp, vp, vp, vp, vp, vp,
                         i32, vp, vp, i32, i32, i32, i32, i32, vp, vp, f32, vp],
    "vae2_ipc_alloc": [i64, C.POINTER(vp), vp],
    "vae2_ipc_open": [vp, C.POINTER(vp)],
    "vae2_ipc_close": [vp],
    "vae2_ipc_free": [vp],
    "vae2_bn_peer_setup": [i32, i32, C.POINTER(vp), vp, vp],
    "vae2_bn_fwd_fused_peer": [vp, vp, vp, vp, i32, i64, i32, i32, i32, i32, i32, vp, vp, vp, vp, vp, f32, f32, vp, vp, vp, vp,
                               i32, i32, i32, i64, i32, vp],
    "vae2_bn_bwd_fused_peer": [vp, vp, vp, vp, vp, vp, i32, i64, i32, i32, i32, i32, i32, i32, i32, vp, vp, vp, vp, vp, vp,
                               i32, vp, vp, i32, i32, i32, i32, i32, f32, i64, i32, vp],
    "vae2_fuse_sum": [C.POINTER(FuseSrc), i32, vp, i32, i32, i32, i32, i32, i32, i32, vp],
    "vae2_fuse_bwd_same": [vp, vp, C.POINTER(FuseDst), i32, i32, i64, i32, i32, i32, i32, vp],
    "vae2_fuse_bwd_up": [vp, vp, vp, i32, i32, i32, i32, i32, i32, i32, i32, i32, i32, i32, i32, vp],
    "vae2_elbo_terms": [vp, i32, vp, i32, vp, vp],
    "vae2_elbo_terms_bwd": [vp, i32, vp],
    "vae2_adam_step": [vp, vp, vp, vp, i64, f32, f32, f32, f32, f32, vp, f32, vp],
    "vae2_clip_u8_to_nchw": [vp, vp, i32, i32, i32, i32, vp],
    "vae2_to_image": [vp, vp, i64, i32, vp],
    "vae2_frame_metrics": [vp, vp, vp, i32, i32, i32, i32, vp],
    "vae2_ssim_level": [vp, vp, vp, i32, i32, i32, i32, f32, vp],
    "vae2_avgpool2": [vp, vp, i32, i32, i32, vp],
}
_PLAIN_INT = {"vae2_abi_version": [], "vae2_bn_max_partials": [], "vae2_elbo_acc_floats": [],
              "vae2_conv2d_tc_supported": [C.POINTER(ConvGeom)], "vae2_conv2d_tf32_supported": [C.POINTER(ConvGeom)]}

EXPORTS = sorted(set(_PROTOS) | set(_PLAIN_INT) | {"vae2_status_string", "vae2_last_cuda_error", "vae2_last_kernel",
                                                     "vae2_conv2d_wgrad_tc_workspace", "vae2_conv2d_wgrad_f32x2_workspace",
                                                     "vae2_conv2d_tf32_dims"})

_lib = None


def lib():
    """Load (once) and return the ctypes handle; RuntimeError if the extension is absent."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                "vae2_b200: native library %s not found. Build it with `python -c 'import __graft_entry__ as g; "
                "g.build()'` (nvcc, sm_100a). There is no CPU or PyTorch fallback for this path." % LIB_PATH)
        h = C.CDLL(LIB_PATH)
        for name, args in _PROTOS.items():
            fn = getattr(h, name)
            fn.argtypes = args
            fn.restype = C.c_int
        for name, args in _PLAIN_INT.items():
            fn = getattr(h, name)
            fn.argtypes = args
            fn.restype = C.c_int
        h.vae2_conv2d_wgrad_tc_workspace.argtypes = [C.POINTER(ConvGeom)]
        h.vae2_conv2d_wgrad_tc_workspace.restype = C.c_longlong
        h.vae2_conv2d_wgrad_f32x2_workspace.argtypes = [C.POINTER(ConvGeom)]
        h.vae2_conv2d_wgrad_f32x2_workspace.restype = C.c_longlong
        h.vae2_conv2d_tf32_dims.argtypes = [C.POINTER(ConvGeom), ip, ip, ip, ip]
        h.vae2_conv2d_tf32_dims.restype = None
        h.vae2_bn_peer_slot_words.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int]
        h.vae2_bn_peer_slot_words.restype = C.c_longlong
        h.vae2_status_string.argtypes = [C.c_int]
        h.vae2_status_string.restype = C.c_char_p
        h.vae2_last_cuda_error.argtypes = []
        h.vae2_last_cuda_error.restype = C.c_char_p
        h.vae2_last_kernel.argtypes = []
        h.vae2_last_kernel.restype = C.c_char_p
        _lib = h
    return _lib


def tf32_dims(geom):
    """(Nf, Kf, NfT, KfT): padded operand extents of the 3xTF32 weight planes for this conv."""
    a, b, c, d = C.c_int(0), C.c_int(0), C.c_int(0), C.c_int(0)
    lib().vae2_conv2d_tf32_dims(C.byref(geom), C.byref(a), C.byref(b), C.byref(c), C.byref(d))
    return a.value, b.value, c.value, d.value


def check(status, what):
    if status != 0:
        h = lib()
        raise RuntimeError("vae2_b200: %s failed: %s (cuda: %s)" % (
            what, h.vae2_status_string(status).decode(), h.vae2_last_cuda_error().decode()))


COUNTERS = {"native_calls": 0}


class _Caller:
    """``call.vae2_bn_apply(...)`` = invoke + status check (+ a count of kernel-launching calls)."""

    def __getattr__(self, name):
        fn = getattr(lib(), name)

        def wrapped(*args):
            COUNTERS["native_calls"] += 1
            check(fn(*args), name)
        wrapped.__name__ = name
        setattr(self, name, wrapped)
        return wrapped


call = _Caller()
