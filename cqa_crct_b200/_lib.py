"""ctypes binding of libcrct_b200.so (include/crct_b200.h).  No fallback: if the library is missing or
the device is not a B200-class GPU, every call raises."""
from __future__ import annotations

import ctypes as C
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, 'libcrct_b200.so')

EPI_BIAS, EPI_BIAS_GELU, EPI_BIAS_RES, EPI_DGELU, EPI_F32 = 0, 1, 2, 3, 4


class CrctError(RuntimeError):
    pass


class GemmArgs(C.Structure):
    _fields_ = [('A', C.c_void_p), ('B', C.c_void_p), ('D', C.c_void_p), ('D2', C.c_void_p),
                ('bias', C.c_void_p), ('aux', C.c_void_p),
                ('M', C.c_int32), ('N', C.c_int32), ('K', C.c_int32),
                ('lda', C.c_int32), ('ldb', C.c_int32), ('ldd', C.c_int32), ('ldaux', C.c_int32),
                ('a_major', C.c_int32), ('b_major', C.c_int32), ('epilogue', C.c_int32),
                ('accumulate', C.c_int32), ('split_k', C.c_int32), ('block_n', C.c_int32),
                ('dropout_p', C.c_float), ('seed', C.c_uint64), ('max_ctas', C.c_int32),
                ('dbg', C.c_int32 * 7)]


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise CrctError(f'{LIB_PATH} not built: run `python -c "import __graft_entry__ as g; g.build()"` '
                            '(there is no CPU / PyTorch fallback for the CRCT hot path)')
        _lib = C.CDLL(LIB_PATH)
        _lib.crct_last_error.restype = C.c_char_p
        for name in EXPORTS:
            getattr(_lib, name)          # fail loudly on a stale library
    return _lib


# every symbol include/crct_b200.h declares (tests check the .so exports exactly these)
EXPORTS = ['crct_last_error', 'crct_version', 'crct_device_check', 'crct_gemm_bf16']


def check(rc: int):
    if rc != 0:
        raise CrctError(f'libcrct_b200 error {rc}: {lib().crct_last_error().decode()}')


def stream_ptr() -> C.c_void_p:
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def ptr(t):
    return None if t is None else C.c_void_p(t.data_ptr())


def gemm(A, B, D, *, M, N, K, a_major=0, b_major=0, epilogue=EPI_BIAS, bias=None, aux=None, D2=None,
         lda=None, ldb=None, ldd=None, ldaux=None, accumulate=0, split_k=0, block_n=0, dropout_p=0.0, seed=0,
         max_ctas=0, dbg=None):
    """D[M,N] = epilogue(A·Bᵀ) — see include/crct_b200.h for operand majors."""
    a = GemmArgs()
    a.A, a.B, a.D, a.D2, a.bias, a.aux = ptr(A), ptr(B), ptr(D), ptr(D2), ptr(bias), ptr(aux)
    a.M, a.N, a.K = M, N, K
    a.lda = lda if lda is not None else A.stride(0)
    a.ldb = ldb if ldb is not None else B.stride(0)
    a.ldd = ldd if ldd is not None else D.stride(0)
    a.ldaux = ldaux if ldaux is not None else (aux.stride(0) if aux is not None else 0)
    a.a_major, a.b_major, a.epilogue = a_major, b_major, epilogue
    a.accumulate, a.split_k, a.block_n = accumulate, split_k, block_n
    a.dropout_p, a.seed, a.max_ctas = dropout_p, seed, max_ctas
    if dbg is not None:
        for i, v in enumerate(dbg):
            a.dbg[i] = v
    check(lib().crct_gemm_bf16(C.byref(a), stream_ptr()))
