"""ctypes binding of the C-ABI library (include/bnn_b200.h) and its in-tree build.

There is no CPU fallback anywhere in this package: if ``libbnn_b200.so`` is missing or the
device is not sm_100 every compute call raises.
"""
import ctypes
import os
import subprocess
import sys

_HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(_HERE, "csrc")
LIB_PATH = os.path.join(CSRC, "libbnn_b200.so")
SOURCES = ["kernels_simt.cu", "kernels_head.cu", "kernels_stats.cu", "kernels_comm.cu", "conv_tc.cu"]
HEADERS = ["common.cuh", "philox.cuh", "../../include/bnn_b200.h"]

F32, F16, BF16, I8 = 0, 1, 2, 3
DROP_NONE, DROP_ELEMENT, DROP_CHANNEL, DROP_MASKSEMBLES = 0, 1, 2, 3
CAL_TOP, CAL_TOP_NORM, CAL_TFP_RESOFTMAX = 0, 1, 2



class DropDesc(ctypes.Structure):
    """struct bnn_drop_desc"""
    _fields_ = [("kind", ctypes.c_int), ("p", ctypes.c_float), ("seed", ctypes.c_uint64),
                ("stream_id", ctypes.c_uint32), ("sample0", ctypes.c_uint32), ("batch", ctypes.c_int),
                ("masks", ctypes.c_void_p), ("n_masks", ctypes.c_int), ("cnt0", ctypes.c_int),
                ("compact_pos", ctypes.c_void_p), ("compact_idx", ctypes.c_void_p), ("compact_c", ctypes.c_int),
                ("nchw_flat", ctypes.c_int)]


OBJ_DIR = os.path.join(CSRC, "build")
COMPILE_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-Xcompiler", "-fPIC"]


def _newer(path, t):
    return os.path.exists(path) and os.path.getmtime(path) > t


def _obj_stale(src, obj):
    if not os.path.exists(obj):
        return True
    t = os.path.getmtime(obj)
    return _newer(src, t) or any(_newer(os.path.join(CSRC, h), t) for h in HEADERS)


def _stale():
    if not os.path.exists(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    return any(_newer(os.path.join(CSRC, f), t) for f in SOURCES + HEADERS)


def build(force=False, verbose=False):
    """Compile every CUDA source for sm_100a into csrc/libbnn_b200.so (nvcc cross-compiles without a GPU): one
    object per source, stale ones recompiled in parallel, then one link.  Returns the library path."""
    if not force and not _stale():
        return LIB_PATH
    nvcc = os.environ.get("NVCC", "nvcc")
    os.makedirs(OBJ_DIR, exist_ok=True)
    jobs, objs = [], []
    for s in SOURCES:
        src, obj = os.path.join(CSRC, s), os.path.join(OBJ_DIR, s[:-3] + ".o")
        if not os.path.exists(src):
            raise RuntimeError("missing CUDA source %s" % src)
        objs.append(obj)
        if force or _obj_stale(src, obj):
            cmd = [nvcc] + COMPILE_FLAGS + ["-c", src, "-o", obj]
            if verbose:
                print(" ".join(cmd), file=sys.stderr)
            jobs.append((cmd, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    for cmd, proc in jobs:
        out, _ = proc.communicate()
        if proc.returncode != 0:
            raise RuntimeError("nvcc failed: %s\n%s" % (" ".join(cmd), out))
    link = [nvcc, "-shared", "-cudart", "static", "-gencode", "arch=compute_100a,code=sm_100a"] + objs + ["-o", LIB_PATH]
    if verbose:
        print(" ".join(link), file=sys.stderr)
    r = subprocess.run(link, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("nvcc link failed:\n%s\n%s" % (r.stdout, r.stderr))
    return LIB_PATH


_SIGS = {
    "bnn_version": (ctypes.c_int, []),
    "bnn_last_error": (ctypes.c_char_p, []),
    "bnn_device_check": (ctypes.c_int, []),
    "bnn_sm_count": (ctypes.c_int, []),
    "bnn_philox_words": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int64, ctypes.c_uint64, ctypes.c_uint32,
                                        ctypes.c_uint32, ctypes.c_void_p]),
    "bnn_philox_keep": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int64, ctypes.c_float, ctypes.c_uint64,
                                       ctypes.c_uint32, ctypes.c_uint32, ctypes.c_void_p]),
    "bnn_nchw_to_nhwc": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p] + [ctypes.c_int] * 5 + [ctypes.c_void_p]),
    "bnn_nchw_to_nhwc_pitch": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p] + [ctypes.c_int] * 6 + [ctypes.c_void_p]),
    "bnn_conv2d_simt": (ctypes.c_int, [ctypes.c_void_p] * 5 + [ctypes.c_int] * 11 +
                        [ctypes.POINTER(DropDesc), ctypes.c_void_p]),
    "bnn_conv2d_tc": (ctypes.c_int, [ctypes.c_void_p] * 5 + [ctypes.c_int] * 9 +
                      [ctypes.POINTER(DropDesc), ctypes.c_void_p]),
    "bnn_conv2d_tc_i8": (ctypes.c_int, [ctypes.c_void_p] * 4 + [ctypes.c_int] * 7 + [ctypes.c_float, ctypes.POINTER(DropDesc),
                                                                       ctypes.c_void_p]),
    "bnn_conv2d_tc_shortcut": (ctypes.c_int, [ctypes.c_void_p] * 5 + [ctypes.c_int] * 9 +
                               [ctypes.POINTER(DropDesc), ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                ctypes.c_void_p]),
    "bnn_conv2d_tc_pooled": (ctypes.c_int, [ctypes.c_void_p] * 5 + [ctypes.c_int] * 9 + [ctypes.c_void_p]),
    "bnn_conv2d_tc_grouped": (ctypes.c_int, [ctypes.c_void_p] * 4 + [ctypes.c_int, ctypes.c_uint32, ctypes.c_uint32] +
                              [ctypes.c_int] * 8 + [ctypes.c_void_p]),
    "bnn_conv2d_tc_gathered": (ctypes.c_int, [ctypes.c_void_p] * 4 + [ctypes.c_int, ctypes.c_uint32, ctypes.c_uint32] +
                               [ctypes.c_int] * 10 + [ctypes.c_uint32, ctypes.c_int, ctypes.c_int, ctypes.c_void_p]),
    "bnn_confidence_exit": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                           ctypes.c_double, ctypes.c_int] + [ctypes.c_void_p] * 4),
    "bnn_top_label": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p] + [ctypes.c_int] * 4 + [ctypes.c_void_p] * 4),
    "bnn_kde_triweight": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_double, ctypes.c_double,
                                         ctypes.c_double, ctypes.c_double, ctypes.c_int, ctypes.c_double, ctypes.c_double,
                                         ctypes.c_void_p, ctypes.c_void_p]),
    "bnn_exit_head_rows": (ctypes.c_int, [ctypes.c_void_p] + [ctypes.c_int] * 7 + [ctypes.c_void_p, ctypes.c_void_p,
                                          ctypes.POINTER(DropDesc)] + [ctypes.c_void_p] * 5 + [ctypes.c_int, ctypes.c_void_p]),
    "bnn_peer_allreduce": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int64, ctypes.POINTER(ctypes.c_void_p), ctypes.c_int,
                                          ctypes.c_int, ctypes.c_int64, ctypes.c_void_p, ctypes.c_void_p]),
    "bnn_peer_allreduce_status": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p]),
    "bnn_dropout": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_int64, ctypes.c_int,
                                   ctypes.c_int, ctypes.c_int, ctypes.POINTER(DropDesc), ctypes.c_void_p]),
    "bnn_channel_affine": (ctypes.c_int, [ctypes.c_void_p] * 4 + [ctypes.c_int, ctypes.c_int64, ctypes.c_int, ctypes.c_int,
                                          ctypes.c_void_p]),
    "bnn_boundary_bits": (ctypes.c_int, [ctypes.c_void_p] * 4 + [ctypes.c_int] * 6 + [ctypes.POINTER(DropDesc), ctypes.c_void_p]),
    "bnn_conv2d_tc_grouped_masked": (ctypes.c_int, [ctypes.c_void_p] * 5 + [ctypes.c_int, ctypes.c_uint32] + [ctypes.c_int] * 7 +
                                     [ctypes.c_void_p]),
    "bnn_conv2d_tc_shortcut_plane": (ctypes.c_int, [ctypes.c_void_p] * 5 + [ctypes.c_int] * 9 +
                                     [ctypes.POINTER(DropDesc), ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p]),
    "bnn_dropout_q8": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_int64, ctypes.c_int,
                                      ctypes.c_int, ctypes.c_int, ctypes.c_float, ctypes.POINTER(DropDesc), ctypes.c_void_p]),
    "bnn_exit_head_q8": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_float] + [ctypes.c_int] * 6 + [ctypes.c_void_p, ctypes.c_void_p,
                                                                                                 ctypes.POINTER(DropDesc)] +
                         [ctypes.c_void_p] * 4 + [ctypes.c_int, ctypes.c_void_p]),
    "bnn_maxpool2d": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p] + [ctypes.c_int] * 6 + [ctypes.c_void_p]),
    "bnn_exit_head": (ctypes.c_int, [ctypes.c_void_p] + [ctypes.c_int] * 7 + [ctypes.c_void_p, ctypes.c_void_p,
                                                                              ctypes.POINTER(DropDesc)] +
                      [ctypes.c_void_p] * 4 + [ctypes.c_int, ctypes.c_void_p]),
    "bnn_exit_head_mma": (ctypes.c_int, [ctypes.c_void_p] + [ctypes.c_int] * 7 + [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
                                                                                  ctypes.POINTER(DropDesc)] +
                          [ctypes.c_void_p] * 4 + [ctypes.c_int, ctypes.c_void_p]),
    "bnn_exit_head_tc": (ctypes.c_int, [ctypes.c_void_p] + [ctypes.c_int] * 7 + [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int,
                                                                                ctypes.c_int, ctypes.POINTER(DropDesc)] +
                         [ctypes.c_void_p] * 6 + [ctypes.c_int, ctypes.c_void_p]),
    "bnn_split16": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64, ctypes.c_int,
                                   ctypes.c_void_p]),
    "bnn_finalize": (ctypes.c_int, [ctypes.c_void_p] * 3 + [ctypes.c_int] * 4 + [ctypes.c_void_p] * 8),
    "bnn_calibration_bins": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_int,
                                            ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
                                            ctypes.c_void_p]),
    "bnn_dataset_metrics": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_int,
                                           ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]),
}

_lib = None


def exported_symbols():
    return sorted(_SIGS)


def load():
    """dlopen the library and declare every signature of include/bnn_b200.h."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            "bayesnn_fpga_b200: %s is missing - run `python -c 'import __graft_entry__ as g; g.build()'` "
            "(there is no CPU or PyTorch fallback for this path)" % LIB_PATH)
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in _SIGS.items():
        fn = getattr(lib, name)       # AttributeError if the .so lacks a declared symbol
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


class BnnError(RuntimeError):
    pass


def check(rc):
    if rc != 0:
        msg = load().bnn_last_error().decode(errors="replace")
        if rc == -1:
            raise ValueError(msg)
        raise BnnError("bnn_b200 error %d: %s" % (rc, msg))


def require_device():
    """Raise unless the current CUDA device is a B200-class (sm_100) part."""
    check(load().bnn_device_check())
