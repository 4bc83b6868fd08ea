"""ctypes binding of libssr_b200.so (the C ABI declared in include/ssr_b200.h).

The library is the product: if it is missing or fails to load this module raises -- there is no
PyTorch / CPU fallback anywhere in the package.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libssr_b200.so")

SSR_NONE, SSR_BF16, SSR_F32, SSR_F32_PLANAR4 = 0, 1, 2, 3
OUT32_NONE, OUT32_NHWC, OUT32_NHWC_ATOMIC, OUT32_NCHW, OUT32_PLANAR4, OUT32_PLANAR4_ACC = 0, 1, 2, 3, 4, 5
PACK_FWD, PACK_DGRAD, PACK_FWD_GEMM, PACK_DGRAD_GEMM, PACK_DGRAD_S2 = 0, 1, 2, 3, 4
PACK_LO = 16   # | PACK_FWD / PACK_DGRAD: the rounding residual w - bf16(w) (split-bf16 tight-parity forward)


class ConvTcArgs(C.Structure):
    """Mirror of struct ssr_conv_tc_args (include/ssr_b200.h)."""
    _fields_ = [
        ("x", C.c_void_p),
        ("n_img", C.c_int32), ("h", C.c_int32), ("w", C.c_int32),
        ("x_pix_stride", C.c_int32),
        ("cin", C.c_int32),
        ("w_packed", C.c_void_p),
        ("r", C.c_int32), ("cout", C.c_int32), ("n_pad", C.c_int32),
        ("bias", C.c_void_p),
        ("act", C.c_int32),
        ("s0", C.c_float),
        ("res1", C.c_void_p), ("res1_kind", C.c_int32), ("res1_pix_stride", C.c_int32), ("s1", C.c_float),
        ("res2", C.c_void_p), ("res2_kind", C.c_int32), ("res2_pix_stride", C.c_int32), ("s2", C.c_float),
        ("mask", C.c_void_p), ("mask_pix_stride", C.c_int32), ("mask_lo", C.c_int32), ("mask_relu", C.c_int32),
        ("out_bf16", C.c_void_p), ("out_pix_stride", C.c_int32),
        ("out_f32", C.c_void_p), ("out32_mode", C.c_int32), ("out32_pix_stride", C.c_int32),
        ("n_tile", C.c_int32), ("mt", C.c_int32), ("splits", C.c_int32), ("res1_cmax", C.c_int32),
        ("out_lo", C.c_int32), ("bias_grad", C.c_void_p), ("bias_grad_scale", C.c_float),
        ("stride", C.c_int32), ("pad_y", C.c_int32), ("pad_x", C.c_int32), ("out_oy", C.c_int32), ("out_ox", C.c_int32),
    ]


class SsrError(RuntimeError):
    pass


_lib = None


def load(build_if_missing=True):
    """Load (building first if the .so is absent and nvcc is present) and return the ctypes handle."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        if not build_if_missing:
            raise SsrError(f"{LIB_PATH} is missing: run `python -m satlas_super_resolution_b200.build`")
        from . import build as _build
        _build.build()
    lib = C.CDLL(LIB_PATH)
    lib.ssr_last_error.restype = C.c_char_p
    lib.ssr_abi_version.restype = C.c_int
    lib.ssr_launch_count.restype = C.c_int64
    lib.ssr_conv_tc.argtypes = [C.POINTER(ConvTcArgs), C.c_void_p]
    lib.ssr_conv_tc.restype = C.c_int
    lib.ssr_conv_tc_chain.argtypes = [C.POINTER(ConvTcArgs), C.c_int32, C.c_void_p]
    lib.ssr_conv_tc_chain.restype = C.c_int
    lib.ssr_conv_tc_chain_acc.argtypes = [C.POINTER(ConvTcArgs), C.c_int32, C.c_void_p]
    lib.ssr_conv_tc_chain_acc.restype = C.c_int
    lib.ssr_conv_tc_chain_acc_supported.argtypes = [C.c_int32, C.c_int32, C.c_int32, C.c_int32]
    lib.ssr_conv_tc_chain_acc_supported.restype = C.c_int
    lib.ssr_rdb_resident_max_blocks.argtypes = [C.c_int32, C.c_int32, C.c_int32]
    lib.ssr_rdb_resident_max_blocks.restype = C.c_int
    lib.ssr_packed_weight_bytes.argtypes = [C.c_int32, C.c_int32, C.c_int32, C.POINTER(C.c_int32)]
    lib.ssr_packed_weight_bytes.restype = C.c_int64
    lib.ssr_pack_conv_weight.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_void_p,
                                         C.c_void_p, C.c_int32, C.c_int32, C.c_void_p]
    lib.ssr_pack_conv_weight.restype = C.c_int
    _bind_optional(lib)
    _lib = lib
    return lib


def _bind_optional(lib):
    """Prototypes of the remaining entry points (declared in ssr_b200.h); bound lazily by name."""
    from . import _protos
    _protos.bind(lib)


def check(rc):
    if rc != 0:
        raise SsrError(f"libssr_b200 error {rc}: {load().ssr_last_error().decode()}")


def launch_count():
    return int(load().ssr_launch_count())
