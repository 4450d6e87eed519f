"""oracle/pyoracle.py -- TEST INFRASTRUCTURE ONLY.

ctypes bindings for
  * libntt_oracle.so  -- our plain-C restatement of the reference hot path (oracle/ntt_oracle.c), and
  * _ref/libntt_ref[_ifma].so -- the reference's own sources compiled by oracle/Makefile.

Only tests/, __graft_entry__.smoke() and bench.py (cpu_baseline / --impl reference) may import this.
The product package never does.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_U64P = np.ctypeslib.ndpointer(dtype=np.uint64, flags="C_CONTIGUOUS")
u64 = C.c_uint64


def build(quiet=True):
    """Compile the checkers (libntt_oracle.so always; _ref/ only where /root/reference exists)."""
    subprocess.run(["make", "-C", _HERE, "all"], check=True,
                   stdout=subprocess.DEVNULL if quiet else None)


def _load(path):
    if not os.path.exists(path):
        return None
    return C.CDLL(path)


class Oracle:
    """The C restatement (always available after build())."""

    def __init__(self):
        path = os.path.join(_HERE, "libntt_oracle.so")
        if not os.path.exists(path):
            build()
        L = C.CDLL(path)
        L.oracle_bitrev.restype = u64
        L.oracle_bitrev.argtypes = [u64, C.c_uint]
        L.oracle_root_table.argtypes = [_U64P, u64, u64, u64]
        L.oracle_shoup_companion.restype = u64
        L.oracle_shoup_companion.argtypes = [u64, u64, C.c_uint]
        L.oracle_shoup_table.argtypes = [_U64P, _U64P, u64, u64, C.c_uint]
        L.oracle_fwd_lazy.argtypes = [_U64P, u64, u64, _U64P, _U64P]
        L.oracle_fwd.argtypes = [_U64P, u64, u64, _U64P, _U64P]
        L.oracle_fwd_dbl.argtypes = [_U64P, _U64P, u64, u64, _U64P, _U64P]
        L.oracle_inv.argtypes = [_U64P, u64, u64, u64, u64, C.c_uint, _U64P, _U64P]
        L.oracle_fwd_batch.argtypes = [_U64P, C.c_size_t, u64, u64, _U64P, _U64P, C.c_uint]
        L.oracle_inv_batch.argtypes = [_U64P, C.c_size_t, u64, u64, u64, u64, _U64P, _U64P, C.c_uint]
        L.oracle_fwd_definition.argtypes = [_U64P, _U64P, u64, u64, u64]
        L.oracle_negacyclic_mul.argtypes = [_U64P, _U64P, _U64P, u64, u64]
        L.oracle_pointwise_mul.argtypes = [_U64P, _U64P, _U64P, u64, u64]
        for f in ("oracle_powmod", "oracle_invmod", "oracle_min_primitive_root_2n", "oracle_fnv1a64"):
            getattr(L, f).restype = u64
        L.oracle_powmod.argtypes = [u64, u64, u64]
        L.oracle_invmod.argtypes = [u64, u64]
        L.oracle_is_prime.argtypes = [u64]
        L.oracle_is_prime.restype = C.c_int
        L.oracle_min_primitive_root_2n.argtypes = [u64, u64]
        L.oracle_fill_uniform.argtypes = [_U64P, C.c_size_t, u64, u64]
        L.oracle_fnv1a64.argtypes = [_U64P, C.c_size_t]
        self.L = L

    # -- tables ------------------------------------------------------------------------------
    def tables(self, N, q, root):
        w = np.empty(N, dtype=np.uint64)
        wc = np.empty(N, dtype=np.uint64)
        self.L.oracle_root_table(w, root, N, q)
        self.L.oracle_shoup_table(wc, w, N, q, 64)
        return w, wc

    def companion(self, v, q, bits=64):
        return int(self.L.oracle_shoup_companion(v, q, bits))

    # -- transforms (return new arrays) ----------------------------------------------------------
    def fwd(self, a, q, w, wc):
        out = np.ascontiguousarray(a, dtype=np.uint64).copy()
        flat = out.reshape(-1, w.shape[0])
        for row in flat:
            self.L.oracle_fwd(row, row.shape[0], q, w, wc)
        return out

    def fwd_lazy(self, a, q, w, wc):
        out = np.ascontiguousarray(a, dtype=np.uint64).copy()
        flat = out.reshape(-1, w.shape[0])
        for row in flat:
            self.L.oracle_fwd_lazy(row, row.shape[0], q, w, wc)
        return out

    def inv(self, a, q, n_inv, wi, wic, n_inv_con=None):
        if n_inv_con is None:
            n_inv_con = self.companion(n_inv, q)
        out = np.ascontiguousarray(a, dtype=np.uint64).copy()
        flat = out.reshape(-1, wi.shape[0])
        for row in flat:
            self.L.oracle_inv(row, row.shape[0], q, n_inv, n_inv_con, 64, wi, wic)
        return out

    def fwd_batch(self, a, q, w, wc, threads=None):
        """oracle_fwd on every row, rows spread over host threads (same arithmetic, see ntt_oracle.c)."""
        out = np.ascontiguousarray(a, dtype=np.uint64).copy()
        N = w.shape[0]
        self.L.oracle_fwd_batch(out.reshape(-1), out.size // N, N, q, w, wc, threads or os.cpu_count() or 1)
        return out

    def inv_batch(self, a, q, n_inv, wi, wic, n_inv_con=None, threads=None):
        if n_inv_con is None:
            n_inv_con = self.companion(n_inv, q)
        out = np.ascontiguousarray(a, dtype=np.uint64).copy()
        N = wi.shape[0]
        self.L.oracle_inv_batch(out.reshape(-1), out.size // N, N, q, n_inv, n_inv_con, wi, wic,
                                threads or os.cpu_count() or 1)
        return out

    def fwd_definition(self, a, q, psi):
        a = np.ascontiguousarray(a, dtype=np.uint64)
        out = np.empty_like(a)
        self.L.oracle_fwd_definition(out, a, a.shape[0], q, psi)
        return out

    def negacyclic_mul(self, a, b, q):
        a = np.ascontiguousarray(a, dtype=np.uint64)
        b = np.ascontiguousarray(b, dtype=np.uint64)
        c = np.empty_like(a)
        self.L.oracle_negacyclic_mul(c, a, b, a.shape[0], q)
        return c

    def pointwise_mul(self, a, b, q):
        a = np.ascontiguousarray(a, dtype=np.uint64).reshape(-1)
        b = np.ascontiguousarray(b, dtype=np.uint64).reshape(-1)
        c = np.empty_like(a)
        self.L.oracle_pointwise_mul(c, a, b, a.shape[0], q)
        return c

    # -- number theory / generators ----------------------------------------------------------------
    def powmod(self, a, e, q):
        return int(self.L.oracle_powmod(a, e, q))

    def invmod(self, a, q):
        return int(self.L.oracle_invmod(a, q))

    def is_prime(self, n):
        return bool(self.L.oracle_is_prime(n))

    def min_root(self, N, q):
        return int(self.L.oracle_min_primitive_root_2n(N, q))

    def uniform(self, n, q, seed):
        a = np.empty(n, dtype=np.uint64)
        self.L.oracle_fill_uniform(a, n, q, seed)
        return a

    def fnv(self, a):
        a = np.ascontiguousarray(a, dtype=np.uint64).reshape(-1)
        return int(self.L.oracle_fnv1a64(a, a.shape[0]))


class Reference:
    """The reference's own code (oracle/_ref/*.so).  `available` is False if it was never built."""

    def __init__(self, want_ifma=True):
        self.ifma = False
        L = None
        if want_ifma and cpu_has_ifma():
            L = _load(os.path.join(_HERE, "_ref", "libntt_ref_ifma.so"))
            self.ifma = L is not None
        if L is None:
            L = _load(os.path.join(_HERE, "_ref", "libntt_ref.so"))
        self.available = L is not None
        self.L = L
        if not self.available:
            return
        L.ref_num_cases.restype = C.c_int
        L.ref_case_params.argtypes = [C.c_int, _U64P]
        L.ref_calc_w.argtypes = [_U64P, u64, u64, u64, u64]
        L.ref_calc_w_con.argtypes = [_U64P, _U64P, u64, u64, u64]
        L.ref_calc_ninv_con.restype = u64
        L.ref_calc_ninv_con.argtypes = [u64, u64, u64]
        L.ref_bit_rev_idx.restype = u64
        L.ref_bit_rev_idx.argtypes = [u64, u64]
        L.ref_expand_w.argtypes = [_U64P, _U64P, u64, u64]
        for f in ("ref_fwd_lazy", "ref_fwd", "ref_fwd_seal", "ref_fwd_radix4", "ref_fwd_radix4x4"):
            getattr(L, f).argtypes = [_U64P, u64, u64, _U64P, _U64P]
        L.ref_fwd_dbl.argtypes = [_U64P, _U64P, u64, u64, _U64P, _U64P]
        for f in ("ref_inv", "ref_inv_seal", "ref_inv_radix4"):
            getattr(L, f).argtypes = [_U64P, u64, u64, u64, u64, _U64P, _U64P]
        L.ref_num_variants.restype = C.c_int
        L.ref_variant_name.restype = C.c_char_p
        L.ref_variant_name.argtypes = [C.c_int]
        L.ref_bench_variant.restype = C.c_double
        L.ref_bench_variant.argtypes = [C.c_int, u64, u64, u64, u64, u64, C.c_int, u64, u64]
        L.ref_run_variant.restype = C.c_int
        L.ref_run_variant.argtypes = [C.c_int, u64, u64, u64, u64, u64, _U64P]

    def cases(self):
        out = []
        buf = np.empty(5, dtype=np.uint64)
        for i in range(self.L.ref_num_cases()):
            self.L.ref_case_params(i, buf)
            m, q, w, w_inv, n_inv = (int(x) for x in buf)
            out.append(dict(idx=i, m=m, q=q, w=w, w_inv=w_inv, n_inv=n_inv))
        return out

    def tables(self, N, q, root):
        m = N.bit_length() - 1
        w = np.empty(N, dtype=np.uint64)
        wc = np.empty(N, dtype=np.uint64)
        self.L.ref_calc_w(w, root, N, q, m)
        self.L.ref_calc_w_con(wc, w, N, q, 64)
        return w, wc

    def ninv_con(self, n_inv, q):
        return int(self.L.ref_calc_ninv_con(n_inv, q, 64))

    def _apply(self, fn, a, N, *args):
        out = np.ascontiguousarray(a, dtype=np.uint64).copy()
        for row in out.reshape(-1, N):
            fn(row, N, *args)
        return out

    def fwd(self, a, q, w, wc):
        return self._apply(self.L.ref_fwd, a, w.shape[0], q, w, wc)

    def fwd_lazy(self, a, q, w, wc):
        return self._apply(self.L.ref_fwd_lazy, a, w.shape[0], q, w, wc)

    def fwd_seal(self, a, q, w, wc):
        return self._apply(self.L.ref_fwd_seal, a, w.shape[0], q, w, wc)

    def inv(self, a, q, n_inv, wi, wic):
        return self._apply(self.L.ref_inv, a, wi.shape[0], q, n_inv, self.ninv_con(n_inv, q), wi, wic)

    def inv_seal(self, a, q, n_inv, wi, wic):
        return self._apply(self.L.ref_inv_seal, a, wi.shape[0], q, n_inv, self.ninv_con(n_inv, q), wi, wic)

    def variants(self):
        return [self.L.ref_variant_name(i).decode() for i in range(self.L.ref_num_variants())]

    def bench(self, variant, m, q, psi, psi_inv, n_inv, threads, calls, seed=1):
        names = self.variants()
        v = names.index(variant)
        return float(self.L.ref_bench_variant(v, m, q, psi, psi_inv, n_inv, threads, calls, seed))

    def run_variant(self, variant, m, q, psi, psi_inv, n_inv, a):
        out = np.ascontiguousarray(a, dtype=np.uint64).copy()
        rc = self.L.ref_run_variant(self.variants().index(variant), m, q, psi, psi_inv, n_inv, out)
        if rc != 0:
            raise RuntimeError("variant %s unavailable" % variant)
        return out


def cpu_has_ifma():
    try:
        with open("/proc/cpuinfo") as f:
            return "avx512ifma" in f.read()
    except OSError:
        return False
