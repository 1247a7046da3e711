"""ctypes binding of libexpertsim_b200.so — the C-ABI boundary of the B200-native hot path.

The prototypes are parsed from ``include/expertsim_b200.h`` so the binding can never drift from the header.
There is NO fallback: if the shared library is missing, or a kernel entry point returns an error, this raises.
PyTorch is used only to own device memory and streams; every pointer handed to the library is ``tensor.data_ptr()``.
"""
from __future__ import annotations

import ctypes
import os
import re
from typing import Dict, List, Tuple

import torch

_PKG_DIR = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_REPO = os.path.dirname(_PKG_DIR)
LIB_PATH = os.path.join(_PKG_DIR, "libexpertsim_b200.so")
HEADER_PATH = os.path.join(_REPO, "include", "expertsim_b200.h")


class ESGroup(ctypes.Structure):
    _fields_ = [("row_start", ctypes.c_int32), ("rows", ctypes.c_int32), ("slot", ctypes.c_int32), ("pass_rows", ctypes.c_int32)]


class ESConvGeom(ctypes.Structure):
    _fields_ = [(n, ctypes.c_int32) for n in ("Hs", "Ws", "C", "Hu", "Wu", "Ho", "Wo", "KH", "KW", "pad", "N")]


class ESConv2d(ctypes.Structure):
    _fields_ = [(n, ctypes.c_int32) for n in ("Ci", "Hi", "Wi", "Co", "Ho", "Wo", "KH", "KW", "stride", "pad")]


class ESTapGeom(ctypes.Structure):
    _fields_ = ([(n, ctypes.c_int32) for n in ("Hs", "Ws", "C", "Hu", "Wu", "Ho", "Wo", "my", "mx", "n_taps")]
                + [("tap_dy", ctypes.c_int8 * 32), ("tap_dx", ctypes.c_int8 * 32), ("tap_koff", ctypes.c_int32 * 32)]
                + [(n, ctypes.c_int32) for n in ("KK", "N", "o_my", "o_oy", "o_mx", "o_ox", "Ho_full", "Wo_full")])


class ESFoldTable(ctypes.Structure):
    _fields_ = [("n_taps", ctypes.c_int32), ("mask", ctypes.c_uint32 * 32)]


def parse_header(path: str = HEADER_PATH) -> Dict[str, Tuple[str, List[Tuple[str, str]]]]:
    """-> {function name: (return type, [(ctype kind, arg name), ...])} for every prototype in the header."""
    src = open(path).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    protos = {}
    for m in re.finditer(r"(?:^|\n)\s*(const char\*|int|void)\s+(es_\w+)\s*\(([^;{}]*?)\)\s*;", src):
        ret, name, args = m.group(1), m.group(2), " ".join(m.group(3).split())
        parsed = []
        if args and args != "void":
            for a in args.split(","):
                a = a.strip()
                aname = re.findall(r"(\w+)$", a)[0]
                if "*" in a:
                    kind = "ptr"
                elif re.search(r"\blong\b", a):
                    kind = "long"
                elif re.search(r"\bdouble\b", a):
                    kind = "double"
                elif re.search(r"\bfloat\b", a):
                    kind = "float"
                elif re.search(r"\b(int|int32_t)\b", a):
                    kind = "int"
                else:
                    raise ValueError(f"cannot classify argument '{a}' of {name}")
                parsed.append((kind, aname))
        protos[name] = (ret, parsed)
    return protos


_CT = {"ptr": ctypes.c_void_p, "long": ctypes.c_long, "double": ctypes.c_double, "float": ctypes.c_float, "int": ctypes.c_int}
_lib = None
_protos = None


def load():
    """Load the shared library (once).  Raises if it has not been built — the product path has no CPU fallback."""
    global _lib, _protos
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(or `make -C generative-dnn-for-physics-simulations-cern_b200/csrc`).  There is no CPU fallback.")
    lib = ctypes.CDLL(LIB_PATH)
    _protos = parse_header()
    for name, (ret, args) in _protos.items():
        fn = getattr(lib, name)  # AttributeError if the library does not export a declared symbol
        fn.argtypes = [_CT[k] for k, _ in args]
        fn.restype = ctypes.c_char_p if ret == "const char*" else ctypes.c_int
    _lib = lib
    return lib


def prototypes():
    load()
    return _protos


def last_error() -> str:
    return load().es_last_error().decode()


def _ptr(x):
    if x is None:
        return None
    if isinstance(x, torch.Tensor):
        if not x.is_cuda:
            raise RuntimeError("libexpertsim_b200 takes device pointers only; got a CPU tensor")
        if not x.is_contiguous():
            raise RuntimeError("non-contiguous tensor passed to the C-ABI")
        return x.data_ptr()
    if isinstance(x, (ctypes.Structure, ctypes.Array)):
        return ctypes.addressof(x)
    if isinstance(x, int):
        return x
    raise TypeError(type(x))


def stream_ptr() -> int:
    return torch.cuda.current_stream().cuda_stream


try:        # raw handle of the current stream without building a torch.cuda.Stream object (called once per launch)
    _raw_stream = torch._C._cuda_getCurrentRawStream
except AttributeError:      # pragma: no cover
    _raw_stream = None

n_calls = 0  # kernels-launching C-ABI calls made so far (bench.py reports it as gpu_launches evidence)
# Optional per-call device timing (bench.py's roofline leg): {"names": set of entry points, "log": []}.  When set, calls to
# the named entry points are bracketed by CUDA events on the launching stream; log rows are (name, args, start, end).
profile = None
_fast = {}   # entry point -> (ctypes function, tuple of argument kinds: 0 pointer, 1 integer, 2 float)
_Tensor = torch.Tensor


def _plan(name):
    lib = load()
    ret, spec = _protos[name]
    kinds = tuple(0 if k == "ptr" else (1 if k in ("int", "long") else 2) for k, _ in spec[:-1])
    _fast[name] = ent = (getattr(lib, name), kinds)
    return ent


def call(name: str, *args):
    """Call ``name(*args, stream)`` on the current torch CUDA stream; tensors are passed as raw device pointers.
    This is the launch path of every kernel (~400 calls per training step), so the per-call host work is kept small:
    the prototype is resolved once per entry point, tensors take a fast path, the stream handle is read raw."""
    global n_calls
    ent = _fast.get(name)
    if ent is None:
        ent = _plan(name)
    fn, kinds = ent
    if len(args) != len(kinds):
        raise TypeError(f"{name} takes {len(kinds)} arguments before the stream, got {len(args)}")
    conv = []
    push = conv.append
    for kind, a in zip(kinds, args):
        if kind == 0:
            if type(a) is _Tensor:
                if not a.is_cuda:
                    raise RuntimeError("libexpertsim_b200 takes device pointers only; got a CPU tensor")
                if not a.is_contiguous():
                    raise RuntimeError("non-contiguous tensor passed to the C-ABI")
                push(a.data_ptr())
            else:
                push(_ptr(a))
        elif kind == 1:
            push(int(a))
        else:
            push(float(a))
    if _raw_stream is not None:
        push(_raw_stream(torch.cuda.current_device()))
    else:
        push(stream_ptr())
    timed = profile is not None and name in profile["names"]
    if timed:
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record()
    rc = fn(*conv)
    if timed:
        ev1.record()
        profile["log"].append((name, tuple(a for a in args if not isinstance(a, torch.Tensor)), ev0, ev1))
    n_calls += 1
    if rc != 0:
        raise RuntimeError(f"{name} failed with status {rc}: {last_error()}")


def device_ok() -> bool:
    return bool(load().es_device_ok())
