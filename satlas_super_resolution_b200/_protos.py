"""ctypes prototypes for the entry points of include/ssr_b200.h beyond the conv core."""
import ctypes as C

i32, f32, vp = C.c_int32, C.c_float, C.c_void_p


class PackDesc(C.Structure):
    """Mirror of struct ssr_pack_desc."""
    _fields_ = [("w", vp), ("dst", vp), ("inv_scale", vp),
                ("cout", i32), ("cin", i32), ("r", i32), ("mode", i32), ("k_pad", i32), ("n_pad", i32)]


PROTOS = {
    "ssr_ingest_nchw": (C.c_int, [vp, i32, vp, i32, i32, i32, i32, i32, i32, f32, vp, vp, vp]),
    "ssr_egress_nchw": (C.c_int, [vp, i32, vp, i32, i32, i32, i32, f32, i32, vp]),
    "ssr_upsample_nearest": (C.c_int, [vp, i32, vp, i32, i32, i32, i32, i32, i32, vp]),
    "ssr_upsample_nearest_bwd": (C.c_int, [vp, i32, vp, i32, i32, i32, i32, i32, i32, vp]),
    "ssr_upsample_bilinear2x": (C.c_int, [vp, i32, vp, i32, i32, i32, i32, i32, vp]),
    "ssr_upsample_bilinear2x_bwd": (C.c_int, [vp, i32, vp, i32, i32, i32, i32, i32, vp]),
    "ssr_pack_conv_weights_batched": (C.c_int, [vp, i32, vp]),
}


def bind(lib):
    for name, (restype, argtypes) in PROTOS.items():
        fn = getattr(lib, name, None)
        if fn is None:
            raise AttributeError(f"libssr_b200.so does not export {name} (stale build?)")
        fn.restype = restype
        fn.argtypes = argtypes
