"""Loaders of the TEST-ONLY native libraries: the CPU oracle (oracle/liborc.so) and the CPU emulation of the
kernels (tests/host_emul/libemul.so). Nothing in the product imports this file."""
import ctypes as C
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def load_oracle(fast=False):
    """liborc.so: -O2 -ffp-contract=off, the bitwise checker. fast=True: liborc_fast.so (-O3 -march=native), the CPU
    baseline timed by bench.py."""
    path = os.path.join(ROOT, "oracle", "liborc_fast.so" if fast else "liborc.so")
    if not os.path.exists(path):
        raise RuntimeError("%s not built: run `make -C oracle` or __graft_entry__.build()" % path)
    lib = C.CDLL(path)
    lib.orc_last_error.restype = C.c_char_p
    return lib


def load_emul():
    path = os.path.join(ROOT, "tests", "host_emul", "libemul.so")
    if not os.path.exists(path):
        raise RuntimeError("tests/host_emul/libemul.so not built: run __graft_entry__.build()")
    lib = C.CDLL(path)
    lib.emul_mpc_create.restype = C.c_void_p
    return lib
