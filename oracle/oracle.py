"""ctypes binding of the CPU oracle (oracle/_build/liblfold_oracle.so).  TEST INFRASTRUCTURE:
only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import this."""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SO = os.path.join(HERE, "_build", "liblfold_oracle.so")
CLI = os.path.join(HERE, "_build", "lfold_oracle")
RLF = os.path.join(HERE, "_ref", "RNALfold")
_lib = None


def build():
    subprocess.run(["make", "-C", HERE, "all"], check=True, stdout=subprocess.DEVNULL)


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(SO):
            build()
        L = C.CDLL(SO)
        L.lfold_oracle_fold.argtypes = [C.c_char_p, C.c_int, C.c_int, C.c_int]
        L.lfold_oracle_fold.restype = C.c_void_p
        L.lfold_oracle_free.argtypes = [C.c_void_p]
        for name in ("nhits", "total", "failed", "W"):
            getattr(L, "lfold_oracle_" + name).argtypes = [C.c_void_p]
            getattr(L, "lfold_oracle_" + name).restype = C.c_int
        for name in ("hit_start", "hit_energy"):
            getattr(L, "lfold_oracle_" + name).argtypes = [C.c_void_p, C.c_int]
            getattr(L, "lfold_oracle_" + name).restype = C.c_int
        L.lfold_oracle_hit_ss.argtypes = [C.c_void_p, C.c_int]
        L.lfold_oracle_hit_ss.restype = C.c_char_p
        L.lfold_oracle_conv.argtypes = [C.c_void_p]
        L.lfold_oracle_conv.restype = C.c_char_p
        for name in ("c", "m", "f3"):
            getattr(L, "lfold_oracle_" + name).argtypes = [C.c_void_p]
            getattr(L, "lfold_oracle_" + name).restype = C.POINTER(C.c_int)
        _lib = L
    return _lib


def fold(seq, span, matrices=False):
    """-> dict(hits=[(ss, energy_dcal, start)], total=int[, c, m, f3])"""
    L = lib()
    b = seq.encode() if isinstance(seq, str) else seq
    r = L.lfold_oracle_fold(b, len(b), span, 1 if matrices else 0)
    try:
        if L.lfold_oracle_failed(r):
            raise RuntimeError("oracle: backtrack failed")
        hits = [(L.lfold_oracle_hit_ss(r, k).decode(), L.lfold_oracle_hit_energy(r, k), L.lfold_oracle_hit_start(r, k))
                for k in range(L.lfold_oracle_nhits(r))]
        out = {"hits": hits, "total": L.lfold_oracle_total(r)}
        if matrices:
            n, W = len(b), L.lfold_oracle_W(r)
            out["c"] = np.ctypeslib.as_array(L.lfold_oracle_c(r), shape=(n + 2, W)).copy()
            out["m"] = np.ctypeslib.as_array(L.lfold_oracle_m(r), shape=(n + 2, W)).copy()
            out["f3"] = np.ctypeslib.as_array(L.lfold_oracle_f3(r), shape=(n + 4,)).copy()
        return out
    finally:
        L.lfold_oracle_free(r)


def fold_text(text, span, binary=None):
    """stdout of the oracle CLI (or of the reference's RNALfold when binary=RLF)."""
    if binary is None:
        if not os.path.exists(CLI):
            build()
        binary = CLI
    return subprocess.run([binary, "-L", str(span)], input=text.encode(), stdout=subprocess.PIPE, check=True).stdout.decode()


def have_rlf():
    return os.path.exists(RLF) and os.access(RLF, os.X_OK)
