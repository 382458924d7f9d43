"""ctypes binding of libstc_b200.so (the C ABI declared in include/stc_b200.h).

Prototypes are parsed from the header so Python argtypes can never drift from the C side.
There is no fallback: if the library is missing, or the device is not cc 10.x, every
call raises.
"""
from __future__ import annotations

import ctypes
import os
import re
from typing import Dict, List, Tuple

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libstc_b200.so")
HEADER_PATH = os.path.join(_HERE, "..", "include", "stc_b200.h")

F32, BF16 = 0, 1
ACT_NONE, ACT_RELU, ACT_HSWISH, ACT_SIGMOID = 0, 1, 2, 3
ENGINE_AUTO, ENGINE_SIMT, ENGINE_TCGEN05 = 0, 1, 2


class GemmDesc(ctypes.Structure):
    _fields_ = [("M", ctypes.c_int), ("N", ctypes.c_int), ("K", ctypes.c_int),
                ("batch1", ctypes.c_int), ("batch2", ctypes.c_int),
                ("sA1", ctypes.c_longlong), ("sA2", ctypes.c_longlong), ("sAm", ctypes.c_longlong), ("sAk", ctypes.c_longlong),
                ("sB1", ctypes.c_longlong), ("sB2", ctypes.c_longlong), ("sBk", ctypes.c_longlong), ("sBn", ctypes.c_longlong),
                ("sC1", ctypes.c_longlong), ("sC2", ctypes.c_longlong), ("sCm", ctypes.c_longlong),
                ("alpha", ctypes.c_float), ("beta", ctypes.c_float)]


_SCALARS = {"int": ctypes.c_int, "long long": ctypes.c_longlong, "float": ctypes.c_float, "double": ctypes.c_double,
            "int64_t": ctypes.c_int64, "unsigned long long": ctypes.c_ulonglong}


def parse_header(path: str = HEADER_PATH) -> Dict[str, Tuple[object, List[object]]]:
    """name -> (restype, [argtypes]) for every `stc_*` prototype in the header."""
    src = open(path).read()
    src = re.sub(r"/\*.*?\*/", " ", src, flags=re.S)
    src = re.sub(r"//[^\n]*", " ", src)
    protos = {}
    for m in re.finditer(r"\b(const\s+char\s*\*|long\s+long|int)\s+(stc_\w+)\s*\(([^;{]*?)\)\s*;", src, flags=re.S):
        ret, name, args = m.group(1), m.group(2), m.group(3).strip()
        restype = ctypes.c_char_p if "char" in ret else (ctypes.c_longlong if "long" in ret else ctypes.c_int)
        argtypes = []
        if args and args != "void":
            for a in args.split(","):
                a = " ".join(a.split())
                if "*" in a:
                    argtypes.append(ctypes.c_void_p)
                else:
                    a = re.sub(r"\bconst\b", "", a).strip()
                    ty = " ".join(a.split()[:-1]) if len(a.split()) > 1 else a
                    if ty not in _SCALARS:
                        raise RuntimeError(f"stc header parse: unknown type '{ty}' in {name}")
                    argtypes.append(_SCALARS[ty])
        protos[name] = (restype, argtypes)
    return protos


class _Lib:
    def __init__(self):
        self._dll = None
        self._protos = None
        self._device_ok = set()

    def load(self):
        if self._dll is not None:
            return self
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} is missing: build it with `python -m stc_unet_b200.build` "
                "(or __graft_entry__.build()). stc_unet_b200 has no CPU / PyTorch fallback.")
        self._dll = ctypes.CDLL(LIB_PATH)
        self._protos = parse_header()
        for name, (restype, argtypes) in self._protos.items():
            fn = getattr(self._dll, name)  # raises AttributeError if a declared symbol is not exported
            fn.restype = restype
            fn.argtypes = argtypes
        return self

    @property
    def symbols(self):
        self.load()
        return sorted(self._protos)

    def last_error(self) -> str:
        return (self._dll.stc_last_error() or b"").decode()

    def ensure_device(self, device_index: int):
        if device_index in self._device_ok:
            return
        self.load()
        if not torch.cuda.is_available():
            raise RuntimeError("stc_unet_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback")
        with torch.cuda.device(device_index):
            rc = self._dll.stc_check_device()
        if rc != 0:
            raise RuntimeError(f"stc_check_device failed: {self.last_error()}")
        self._device_ok.add(device_index)

    def raw(self, name):
        self.load()
        return getattr(self._dll, name)

    def call(self, name: str, *args):
        """Calls an int-returning entry point; tensors become device pointers, None becomes NULL."""
        self.load()
        fn = getattr(self._dll, name)
        conv = []
        for a in args:
            if isinstance(a, torch.Tensor):
                conv.append(a.data_ptr())
            elif isinstance(a, ctypes.Structure):
                conv.append(ctypes.addressof(a))
            else:
                conv.append(a)
        rc = fn(*conv)
        if rc != 0:
            raise RuntimeError(f"{name} failed (code {rc}): {self.last_error()}")


lib = _Lib()


def stream_ptr() -> int:
    return torch.cuda.current_stream().cuda_stream


def dtype_code(dt: torch.dtype) -> int:
    if dt == torch.float32:
        return F32
    if dt == torch.bfloat16:
        return BF16
    raise TypeError(f"stc_unet_b200 supports float32 and bfloat16 activations, got {dt}")
