"""ctypes prototypes for the entry points of include/ssr_b200.h beyond the conv core."""
import ctypes as C


def bind(lib):
    for name, (restype, argtypes) in PROTOS.items():
        fn = getattr(lib, name, None)
        if fn is None:
            raise AttributeError(f"libssr_b200.so does not export {name} (stale build?)")
        fn.restype = restype
        fn.argtypes = argtypes


PROTOS = {}
