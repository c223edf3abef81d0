"""ctypes binding of libmft_b200.so (C ABI declared in include/mft_b200.h; test / tuning hooks in
mft_b200/csrc/mft_b200_internal.h).

The product path has no CPU fallback: if the shared library is missing or fails to load, every
entry point raises.  Loading the library does not need a GPU (symbol checks run on CPU)."""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, 'libmft_b200.so')

NUM_LAYERS = 47
MAX_PAIRS = 8

_lib = None


class MftB200Error(RuntimeError):
    pass


def _declare(lib):
    vp, ci, cf = C.c_void_p, C.c_int, C.c_float
    sig = {
        'mftb200_create': (ci, [C.POINTER(vp)]),
        'mftb200_destroy': (None, [vp]),
        'mftb200_last_error': (C.c_char_p, [vp]),
        'mftb200_version': (C.c_char_p, []),
        'mftb200_upload_layer': (ci, [vp, ci, vp, vp, ci, ci, ci]),
        'mftb200_configure': (ci, [vp, ci, ci, ci, ci, ci]),
        'mftb200_encode_frame': (ci, [vp, vp, ci, ci, vp]),
        'mftb200_is_pinned_host': (ci, [vp]),
        'mftb200_slot_buffers': (ci, [vp, C.POINTER(vp), C.POINTER(vp), C.POINTER(vp), C.POINTER(C.c_size_t), vp]),
        'mftb200_raft_refine': (ci, [vp, ci, C.POINTER(ci), C.POINTER(ci), vp, vp]),
        'mftb200_raft_refine_init': (ci, [vp, ci, C.POINTER(ci), C.POINTER(ci), vp, vp, vp]),
        'mftb200_chain_select': (ci, [ci, C.POINTER(vp), vp, cf, ci, ci, vp, vp, vp]),
        'mftb200_warp_backward': (ci, [vp, vp, ci, ci, ci, ci, vp, vp]),
        'mftb200_sample_points': (ci, [vp, ci, ci, ci, vp, ci, ci, vp, vp]),
        'mftb200_warp_forward': (ci, [vp, vp, vp, ci, ci, ci, ci, cf, vp, vp, vp]),
        'mftb200_device_error_flag': (ci, [vp]),
        'mftb200_error_flag_async': (ci, [vp, vp]),
        'mftb200_error_flag_poll': (ci, [vp]),
        'mftb200_wait_frame_copied': (ci, [vp]),
        'mftb200_set_option': (ci, [vp, C.c_char_p, ci]),
        'mftb200_set_global_option': (ci, [C.c_char_p, ci]),
        'mftb200_debug_buffer': (ci, [vp, C.c_char_p, C.POINTER(vp), C.POINTER(C.c_size_t)]),
        'mftb200_debug_read': (ci, [vp, C.c_char_p, vp, C.c_size_t]),
        'mftb200_profile_fetch': (ci, [vp, C.POINTER(C.c_double), C.POINTER(C.c_longlong)]),
        'mftb200_profile_steps': (ci, [vp, C.POINTER(C.c_float), C.POINTER(ci), ci, C.POINTER(ci)]),
        'mftb200_launch_count': (C.c_longlong, [vp]),
        'mftb200_conv2d_bench': (ci, [vp, ci, ci, ci, ci, ci, vp, vp, ci, ci, ci, ci, ci, ci, vp, ci, ci, ci, ci,
                                      C.POINTER(C.c_float), vp]),
        'mftb200_conv2d_bench2': (ci, [vp, ci, ci, ci, ci, ci, vp, vp, ci, ci, ci, ci, ci, ci, vp, ci, ci, ci, ci,
                                       C.POINTER(C.c_float), vp, vp]),
        'mftb200_conv2d_test': (ci, [vp, ci, ci, ci, ci, ci, vp, vp, ci, ci, ci, ci, ci, ci, vp, ci, vp]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(lib, name)          # AttributeError if the symbol is missing
        fn.restype = res
        fn.argtypes = args
    return sorted(sig)


def lib():
    global _lib
    if _lib is None:
        if not os.path.isfile(LIB_PATH):
            raise MftB200Error(f'{LIB_PATH} not built: run `python -m mft_b200.build` (there is no CPU fallback)')
        handle = C.CDLL(LIB_PATH)
        _declare(handle)
        _lib = handle
    return _lib


def exported_symbols():
    return _declare(C.CDLL(LIB_PATH))


def check(code, ctx=None):
    if code != 0:
        msg = lib().mftb200_last_error(ctx)
        raise MftB200Error(f'mft_b200 error {code}: {msg.decode() if msg else "?"}')
