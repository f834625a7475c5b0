"""ctypes binding of libpmd_sm100.so (the C ABI declared in include/pmd_sm100.h).

There is no CPU fallback: if the shared library is missing the import of any compute entry point
fails loudly with instructions to build it."""
import ctypes
import os
import re

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("PMD_LIB_PATH") or os.path.join(HERE, "libpmd_sm100.so")   # override: debug builds of the library
HEADER_PATH = os.path.join(HERE, "..", "include", "pmd_sm100.h")

_lib = None

c_i64, c_int, c_f32, c_vp = ctypes.c_int64, ctypes.c_int, ctypes.c_float, ctypes.c_void_p

_KIND = {"p": c_vp, "l": c_i64, "i": c_int, "f": c_f32}


def _parse_header():
    """Derive the ctypes signature of every `int pmd_*(...)` entry point from the header itself, so
    the binding cannot drift from the declared ABI.  Kinds: p pointer, l int64_t, i int, f float."""
    with open(HEADER_PATH) as f:
        text = re.sub(r"/\*.*?\*/", "", f.read(), flags=re.S)
    sigs = {}
    for m in re.finditer(r"int\s+(pmd_[a-z0-9_]+)\s*\((.*?)\);", text, flags=re.S):
        name, args = m.group(1), m.group(2).strip()
        kinds = ""
        if args != "void":
            for a in args.split(","):
                a = a.strip()
                if "*" in a:
                    kinds += "p"
                elif a.startswith("int64_t"):
                    kinds += "l"
                elif a.startswith("int "):
                    kinds += "i"
                elif a.startswith("float "):
                    kinds += "f"
                else:
                    raise RuntimeError("unparsed argument %r in %s" % (a, name))
        sigs[name] = kinds
    return sigs


def declared_symbols():
    """Every function name declared in include/pmd_sm100.h (used by the symbol-export test)."""
    with open(HEADER_PATH) as f:
        text = f.read()
    return sorted(set(re.findall(r"\b(pmd_[a-z0-9_]+)\s*\(", text)))


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                "libpmd_sm100.so is missing (%s). Build it with `python -m localmd_b200._build` "
                "(nvcc, sm_100a). There is no CPU fallback." % LIB_PATH
            )
        L = ctypes.CDLL(LIB_PATH)
        L.pmd_last_error.restype = ctypes.c_char_p
        L.pmd_last_error.argtypes = []
        L.pmd_abi_version.restype = c_int
        L.pmd_block_project_ts_workspace_bytes.restype = c_i64   # the one entry point that returns a size, not a status
        L.pmd_block_project_ts_workspace_bytes.argtypes = [c_i64, c_i64, c_i64]
        for name, sig in _parse_header().items():
            fn = getattr(L, name)
            fn.restype = c_int
            fn.argtypes = [_KIND[k] for k in sig]
        _lib = L
    return _lib


class PMDKernelError(RuntimeError):
    pass


def check(rc, name=""):
    if rc != 0:
        msg = lib().pmd_last_error().decode("utf-8", "replace")
        raise PMDKernelError("%s failed (code %d): %s" % (name, rc, msg))
