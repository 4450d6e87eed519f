"""B200-native negacyclic NTT -- Python view of the C-ABI in include/ntt_b200.h.

The product is the shared library ``libntt_b200.so`` (C host code + hand-written sm_100a kernels) built
in this directory by ``make``.  This module only binds it with ctypes so tests, ``bench.py`` and
``torch.distributed`` plumbing can drive it; PyTorch is used for device memory and streams, never for
arithmetic.  There is no fallback: if the library is missing the import raises, and every call raises
``NttError`` when the library reports NTT_B200_ERROR (e.g. no CUDA device).

Entry points mirror the reference's operator interface (include/ntt_reference.h:13-65):
``fwd_ntt_ref_harvey``, ``fwd_ntt_ref_harvey_lazy``, ``inv_ntt_ref_harvey``, ``fwd_ntt_ref_harvey_dbl``
on host arrays, plus the plan/batch API for device-resident data.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libntt_b200.so")
DROPIN_PATH = os.path.join(_HERE, "libntt_b200_dropin.so")

u64 = C.c_uint64
_U64P = np.ctypeslib.ndpointer(dtype=np.uint64, flags="C_CONTIGUOUS")


class NttError(RuntimeError):
    pass


def build(verbose=False):
    """Compile libntt_b200.so / libntt_b200_dropin.so in-tree (nvcc -gencode arch=compute_100a,code=sm_100a)."""
    subprocess.run(["make", "-j%d" % (os.cpu_count() or 4), "-C", _HERE, "all"], check=True,
                   stdout=None if verbose else subprocess.DEVNULL)


def _bind():
    if not os.path.exists(LIB_PATH):
        raise ImportError("%s not built: run `make -C %s` (needs nvcc); there is no CPU fallback" % (LIB_PATH, _HERE))
    L = C.CDLL(LIB_PATH)
    vp, sz, i = C.c_void_p, C.c_size_t, C.c_int
    L.ntt_b200_last_error.restype = C.c_char_p
    L.ntt_b200_version.restype = C.c_char_p
    L.ntt_b200_device_count.restype = i
    L.ntt_b200_configure.argtypes = [C.c_char_p, i]
    L.ntt_b200_plan_create.argtypes = [C.POINTER(vp), i, u64, u64, vp, vp, vp, vp, u64, u64]
    L.ntt_b200_plan_create_psi.argtypes = [C.POINTER(vp), i, u64, u64, u64]
    L.ntt_b200_plan_destroy.argtypes = [vp]
    for f in ("ntt_b200_plan_n", "ntt_b200_plan_q"):
        getattr(L, f).restype = u64
        getattr(L, f).argtypes = [vp]
    L.ntt_b200_plan_device.argtypes = [vp]
    L.ntt_b200_plan_is_lazy.argtypes = [vp]
    L.ntt_b200_plan_describe.argtypes = [vp, i, C.c_char_p, sz, C.POINTER(i)]
    L.ntt_b200_plan_export_tables.argtypes = [vp, vp, vp, vp, vp, C.POINTER(u64), C.POINTER(u64)]
    L.ntt_b200_unordered_index.restype = u64
    L.ntt_b200_unordered_index.argtypes = [vp, u64]
    for f in ("ntt_b200_fwd_batch", "ntt_b200_fwd_lazy_batch", "ntt_b200_inv_batch", "ntt_b200_fwd_unordered_batch",
              "ntt_b200_inv_unordered_batch"):
        getattr(L, f).argtypes = [vp, vp, sz, vp]
    for f in ("ntt_b200_fwd_rns", "ntt_b200_inv_rns"):
        getattr(L, f).argtypes = [C.POINTER(vp), sz, vp, sz, vp]
    for f in ("ntt_b200_fwd_tail_block", "ntt_b200_inv_tail_block"):
        getattr(L, f).argtypes = [vp, vp, C.c_uint32, C.c_uint32, vp]
    for f in ("ntt_b200_fwd_tail_gather", "ntt_b200_inv_tail_scatter"):
        getattr(L, f).argtypes = [vp, C.POINTER(vp), vp, C.c_uint32, C.c_uint32, vp]
    for f in ("ntt_b200_fwd_tail_gather_batch", "ntt_b200_inv_tail_scatter_batch"):
        getattr(L, f).argtypes = [vp, C.POINTER(vp), vp, C.c_uint32, C.c_uint32, sz, vp]
    L.ntt_b200_peer_barrier.argtypes = [i, C.POINTER(vp), vp, C.c_uint32, C.c_uint32, C.c_uint32, vp, vp]
    L.ntt_b200_ipc_export.argtypes = [i, vp, C.c_char_p]
    L.ntt_b200_ipc_open.argtypes = [i, C.c_char_p, C.POINTER(vp)]
    L.ntt_b200_ipc_close.argtypes = [i, vp]
    L.ntt_b200_plan_set_inverse_scale.argtypes = [vp, u64]
    L.ntt_b200_negacyclic_mul_batch.argtypes = [vp, vp, vp, vp, sz, vp]
    L.ntt_b200_pointwise_mul_batch.argtypes = [vp, vp, vp, vp, sz, vp]
    for f in ("ntt_b200_fwd_batch_host", "ntt_b200_inv_batch_host"):
        getattr(L, f).argtypes = [vp, vp, sz]
    L.ntt_b200_fwd_mul_inv_batch.argtypes = [vp, vp, vp, sz, vp]
    L.ntt_b200_fwd_mul_inv_batch_host.argtypes = [vp, vp, vp, sz]
    L.ntt_b200_multi_create.argtypes = [C.POINTER(vp), C.POINTER(i), i, u64, u64, u64]
    L.ntt_b200_multi_destroy.argtypes = [vp]
    L.ntt_b200_multi_devices.argtypes = [vp]
    L.ntt_b200_multi_plan.restype = vp
    L.ntt_b200_multi_plan.argtypes = [vp, i]
    L.ntt_b200_multi_last_error.restype = C.c_char_p
    L.ntt_b200_shard_range.restype = None
    L.ntt_b200_shard_range.argtypes = [sz, i, i, C.POINTER(sz), C.POINTER(sz)]
    for f in ("ntt_b200_multi_fwd_batch", "ntt_b200_multi_inv_batch"):
        getattr(L, f).argtypes = [vp, C.POINTER(vp), C.POINTER(sz), C.POINTER(vp)]
    L.ntt_b200_multi_sync.argtypes = [vp]
    for f in ("ntt_b200_multi_fwd_batch_host", "ntt_b200_multi_inv_batch_host"):
        getattr(L, f).argtypes = [vp, vp, sz]
    L.ntt_b200_multi_fwd_mul_inv_batch_host.argtypes = [vp, vp, C.POINTER(vp), sz]
    for f in ("ntt_b200_fwd_rns_multi", "ntt_b200_inv_rns_multi"):
        getattr(L, f).argtypes = [C.POINTER(vp), sz, C.POINTER(vp), sz]
    L.ntt_b200_host_alloc.argtypes = [C.POINTER(vp), sz]
    L.ntt_b200_host_free.argtypes = [vp]
    L.ntt_b200_device_alloc.argtypes = [i, C.POINTER(vp), sz]
    L.ntt_b200_device_free.argtypes = [i, vp]
    L.ntt_b200_memcpy_h2d.argtypes = [i, vp, vp, sz]
    L.ntt_b200_memcpy_d2h.argtypes = [i, vp, vp, sz]
    L.ntt_b200_device_sync.argtypes = [i]
    L.ntt_b200_fwd_ntt_ref_harvey_lazy.argtypes = [_U64P, u64, u64, _U64P, _U64P]
    L.ntt_b200_fwd_ntt_ref_harvey.argtypes = [_U64P, u64, u64, _U64P, _U64P]
    L.ntt_b200_inv_ntt_ref_harvey.argtypes = [_U64P, u64, u64, u64, u64, u64, _U64P, _U64P]
    L.ntt_b200_fwd_ntt_ref_harvey_dbl.argtypes = [_U64P, _U64P, u64, u64, _U64P, _U64P]
    L.ntt_b200_dropin_reset.restype = None
    L.ntt_b200_bit_rev_idx.restype = u64
    L.ntt_b200_bit_rev_idx.argtypes = [u64, u64]
    L.ntt_b200_calc_w.argtypes = [_U64P, u64, u64, u64]
    L.ntt_b200_calc_w_con.argtypes = [_U64P, _U64P, u64, u64, u64]
    for f in ("ntt_b200_calc_ninv_con", "ntt_b200_pow_mod"):
        getattr(L, f).restype = u64
        getattr(L, f).argtypes = [u64, u64, u64]
    for f in ("ntt_b200_inv_mod", "ntt_b200_min_primitive_root"):
        getattr(L, f).restype = u64
        getattr(L, f).argtypes = [u64, u64]
    L.ntt_b200_is_prime.argtypes = [u64]
    return L


lib = _bind()

#: every symbol include/ntt_b200.h declares (checked against the header and the .so by the CPU tests)
EXPORTS = [
    "ntt_b200_last_error", "ntt_b200_device_count", "ntt_b200_version", "ntt_b200_configure",
    "ntt_b200_plan_create", "ntt_b200_plan_create_psi", "ntt_b200_plan_destroy",
    "ntt_b200_plan_n", "ntt_b200_plan_q", "ntt_b200_plan_device", "ntt_b200_plan_is_lazy",
    "ntt_b200_plan_export_tables", "ntt_b200_plan_describe",
    "ntt_b200_fwd_batch", "ntt_b200_fwd_lazy_batch", "ntt_b200_inv_batch",
    "ntt_b200_fwd_unordered_batch", "ntt_b200_inv_unordered_batch", "ntt_b200_unordered_index",
    "ntt_b200_fwd_rns", "ntt_b200_inv_rns",
    "ntt_b200_fwd_tail_block", "ntt_b200_inv_tail_block", "ntt_b200_plan_set_inverse_scale",
    "ntt_b200_fwd_tail_gather", "ntt_b200_inv_tail_scatter", "ntt_b200_peer_barrier",
    "ntt_b200_fwd_tail_gather_batch", "ntt_b200_inv_tail_scatter_batch",
    "ntt_b200_ipc_export", "ntt_b200_ipc_open", "ntt_b200_ipc_close",
    "ntt_b200_negacyclic_mul_batch", "ntt_b200_pointwise_mul_batch",
    "ntt_b200_fwd_batch_host", "ntt_b200_inv_batch_host",
    "ntt_b200_fwd_mul_inv_batch", "ntt_b200_fwd_mul_inv_batch_host",
    "ntt_b200_multi_create", "ntt_b200_multi_destroy", "ntt_b200_multi_devices", "ntt_b200_multi_plan",
    "ntt_b200_multi_last_error", "ntt_b200_shard_range",
    "ntt_b200_multi_fwd_batch", "ntt_b200_multi_inv_batch", "ntt_b200_multi_sync",
    "ntt_b200_multi_fwd_batch_host", "ntt_b200_multi_inv_batch_host", "ntt_b200_multi_fwd_mul_inv_batch_host",
    "ntt_b200_fwd_rns_multi", "ntt_b200_inv_rns_multi",
    "ntt_b200_host_alloc", "ntt_b200_host_free", "ntt_b200_device_alloc", "ntt_b200_device_free",
    "ntt_b200_memcpy_h2d", "ntt_b200_memcpy_d2h", "ntt_b200_device_sync",
    "ntt_b200_fwd_ntt_ref_harvey_lazy", "ntt_b200_fwd_ntt_ref_harvey", "ntt_b200_inv_ntt_ref_harvey",
    "ntt_b200_fwd_ntt_ref_harvey_dbl", "ntt_b200_dropin_reset",
    "ntt_b200_bit_rev_idx", "ntt_b200_calc_w", "ntt_b200_calc_w_con", "ntt_b200_calc_ninv_con",
    "ntt_b200_pow_mod", "ntt_b200_inv_mod", "ntt_b200_is_prime", "ntt_b200_min_primitive_root",
]


def _check(rc, what):
    if rc != 0:
        raise NttError("%s: %s" % (what, lib.ntt_b200_last_error().decode()))


def device_count():
    return int(lib.ntt_b200_device_count())


def version():
    return lib.ntt_b200_version().decode()


# ---- raw device buffers and CUDA IPC (peer-memory exchange of the distributed transform) -------------------

def device_alloc(device, nbytes):
    p = C.c_void_p()
    _check(lib.ntt_b200_device_alloc(device, C.byref(p), nbytes), "device_alloc")
    return p.value


def device_free(device, ptr):
    _check(lib.ntt_b200_device_free(device, ptr), "device_free")


def memcpy_h2d(device, d_ptr, h_arr):
    _check(lib.ntt_b200_memcpy_h2d(device, d_ptr, h_arr.ctypes.data, h_arr.nbytes), "memcpy_h2d")


def memcpy_d2h(device, h_arr, d_ptr):
    _check(lib.ntt_b200_memcpy_d2h(device, h_arr.ctypes.data, d_ptr, h_arr.nbytes), "memcpy_d2h")


def device_sync(device):
    _check(lib.ntt_b200_device_sync(device), "device_sync")


def ipc_export(device, d_ptr):
    h = C.create_string_buffer(64)
    _check(lib.ntt_b200_ipc_export(device, d_ptr, h), "ipc_export")
    return h.raw


def ipc_open(device, handle):
    p = C.c_void_p()
    _check(lib.ntt_b200_ipc_open(device, handle, C.byref(p)), "ipc_open")
    return p.value


def ipc_close(device, d_ptr):
    _check(lib.ntt_b200_ipc_close(device, d_ptr), "ipc_close")


def peer_barrier(device, peer_flags, my_flags, rank, world, epoch, d_timed_out, stream=None):
    _check(lib.ntt_b200_peer_barrier(device, peer_flags, my_flags, rank, world, epoch, d_timed_out,
                                     _stream_ptr(stream)), "peer_barrier")


def configure(key, value):
    """Kernel selection for benchmarks / A-B tests: configure("fp64", 0), configure("ring", 0)."""
    _check(lib.ntt_b200_configure(key.encode(), int(value)), "configure")


def _ptr(x):
    """Raw address of a torch tensor (device or host), a numpy array, an int, or None."""
    if x is None:
        return None
    if isinstance(x, int):
        return x
    if isinstance(x, np.ndarray):
        assert x.dtype == np.uint64 and x.flags["C_CONTIGUOUS"]
        return x.ctypes.data
    return x.data_ptr()  # torch tensor (int64 storage viewed as uint64 words)


def _stream_ptr(stream):
    if stream is None:
        return None
    return getattr(stream, "cuda_stream", stream)


# ---- host-side table builders (C replacements of include/internal/pre_compute.h) -------------------

def calc_w(root, N, q):
    out = np.empty(N, dtype=np.uint64)
    _check(lib.ntt_b200_calc_w(out, root, N, q), "calc_w")
    return out


def calc_w_con(w, q, word_size=64):
    w = np.ascontiguousarray(w, dtype=np.uint64)
    out = np.empty_like(w)
    _check(lib.ntt_b200_calc_w_con(out, w, w.shape[0], q, word_size), "calc_w_con")
    return out


def calc_ninv_con(n_inv, q, word_size=64):
    return int(lib.ntt_b200_calc_ninv_con(n_inv, q, word_size))


def pow_mod(a, e, q):
    return int(lib.ntt_b200_pow_mod(a, e, q))


def inv_mod(a, q):
    return int(lib.ntt_b200_inv_mod(a, q))


def is_prime(n):
    return bool(lib.ntt_b200_is_prime(n))


def min_primitive_root(N, q):
    return int(lib.ntt_b200_min_primitive_root(N, q))


# ---- plans -------------------------------------------------------------------------------------------

class Plan:
    """One (device, N, q, psi): owns the device twiddle tables.  See ntt_b200_plan_create[_psi]."""

    def __init__(self, handle, N, q):
        self._h = handle
        self.N, self.q = N, q

    @classmethod
    def from_tables(cls, N, q, w=None, w_con=None, w_inv=None, w_inv_con=None, n_inv=0, n_inv_con=0, device=0):
        keep = [None if t is None else np.ascontiguousarray(t, dtype=np.uint64) for t in (w, w_con, w_inv, w_inv_con)]
        h = C.c_void_p()
        _check(lib.ntt_b200_plan_create(C.byref(h), device, N, q, *[_ptr(t) for t in keep], n_inv, n_inv_con),
               "plan_create")
        return cls(h, N, q)

    @classmethod
    def from_psi(cls, N, q, psi, device=0):
        h = C.c_void_p()
        _check(lib.ntt_b200_plan_create_psi(C.byref(h), device, N, q, psi), "plan_create_psi")
        return cls(h, N, q)

    def close(self):
        if self._h:
            lib.ntt_b200_plan_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def is_lazy(self):
        return bool(lib.ntt_b200_plan_is_lazy(self._h))

    @property
    def device(self):
        return int(lib.ntt_b200_plan_device(self._h))

    def describe(self, inverse=False):
        """(kernel names, launches) of one transform, e.g. ("k_ring_fp<14,fwd>", 1)."""
        buf, n = C.create_string_buffer(256), C.c_int()
        _check(lib.ntt_b200_plan_describe(self._h, int(inverse), buf, 256, C.byref(n)), "plan_describe")
        return buf.value.decode(), int(n.value)

    def export_tables(self, inverse=True):
        w = np.empty(self.N, dtype=np.uint64)
        wc = np.empty(self.N, dtype=np.uint64)
        wi = np.empty(self.N, dtype=np.uint64) if inverse else None
        wic = np.empty(self.N, dtype=np.uint64) if inverse else None
        n_inv, n_inv_con = u64(), u64()
        _check(lib.ntt_b200_plan_export_tables(self._h, _ptr(w), _ptr(wc), _ptr(wi), _ptr(wic), C.byref(n_inv),
                                               C.byref(n_inv_con)), "export_tables")
        return dict(w=w, w_con=wc, w_inv=wi, w_inv_con=wic, n_inv=int(n_inv.value), n_inv_con=int(n_inv_con.value))

    # device-resident data: `d_a` is a torch cuda tensor (int64 storage) or a raw device address
    def fwd(self, d_a, batch, stream=None):
        _check(lib.ntt_b200_fwd_batch(self._h, _ptr(d_a), batch, _stream_ptr(stream)), "fwd_batch")

    def fwd_lazy(self, d_a, batch, stream=None):
        _check(lib.ntt_b200_fwd_lazy_batch(self._h, _ptr(d_a), batch, _stream_ptr(stream)), "fwd_lazy_batch")

    def inv(self, d_a, batch, stream=None):
        _check(lib.ntt_b200_inv_batch(self._h, _ptr(d_a), batch, _stream_ptr(stream)), "inv_batch")

    # order-agnostic variants (pointwise consumers); unordered_index(i) = position in the reference's output
    def fwd_unordered(self, d_a, batch, stream=None):
        _check(lib.ntt_b200_fwd_unordered_batch(self._h, _ptr(d_a), batch, _stream_ptr(stream)), "fwd_unordered_batch")

    def inv_unordered(self, d_a, batch, stream=None):
        _check(lib.ntt_b200_inv_unordered_batch(self._h, _ptr(d_a), batch, _stream_ptr(stream)), "inv_unordered_batch")

    def unordered_index(self, i):
        return int(lib.ntt_b200_unordered_index(self._h, i))

    def negacyclic_mul(self, d_c, d_a, d_b, batch, stream=None):
        _check(lib.ntt_b200_negacyclic_mul_batch(self._h, _ptr(d_c), _ptr(d_a), _ptr(d_b), batch,
                                                 _stream_ptr(stream)), "negacyclic_mul_batch")

    def pointwise_mul(self, d_c, d_a, d_b, batch, stream=None):
        _check(lib.ntt_b200_pointwise_mul_batch(self._h, _ptr(d_c), _ptr(d_a), _ptr(d_b), batch,
                                                _stream_ptr(stream)), "pointwise_mul_batch")

    # pieces of a transform spread over 2^log2_parts GPUs (see fourstep.py)
    def fwd_tail_block(self, d_block, log2_parts, block, stream=None):
        _check(lib.ntt_b200_fwd_tail_block(self._h, _ptr(d_block), log2_parts, block, _stream_ptr(stream)),
               "fwd_tail_block")

    def inv_tail_block(self, d_block, log2_parts, block, stream=None):
        _check(lib.ntt_b200_inv_tail_block(self._h, _ptr(d_block), log2_parts, block, _stream_ptr(stream)),
               "inv_tail_block")

    # the same, fused with the exchange over peer memory: peer_slices = ctypes array of G device pointers
    def fwd_tail_gather(self, peer_slices, d_block, log2_parts, rank, stream=None, batch=1):
        _check(lib.ntt_b200_fwd_tail_gather_batch(self._h, peer_slices, _ptr(d_block), log2_parts, rank, batch,
                                                  _stream_ptr(stream)), "fwd_tail_gather")

    def inv_tail_scatter(self, peer_slices, d_block, log2_parts, rank, stream=None, batch=1):
        _check(lib.ntt_b200_inv_tail_scatter_batch(self._h, peer_slices, _ptr(d_block), log2_parts, rank, batch,
                                                   _stream_ptr(stream)), "inv_tail_scatter")

    def set_inverse_scale(self, scale):
        _check(lib.ntt_b200_plan_set_inverse_scale(self._h, scale), "plan_set_inverse_scale")

    # host-resident data: numpy uint64 array or pinned torch tensor, transformed in place
    def fwd_host(self, h_a, batch):
        _check(lib.ntt_b200_fwd_batch_host(self._h, _ptr(h_a), batch), "fwd_batch_host")

    def inv_host(self, h_a, batch):
        _check(lib.ntt_b200_inv_batch_host(self._h, _ptr(h_a), batch), "inv_batch_host")

    # a[b] <- INTT(NTT(a[b]) .* m): d_m = N canonical residues on the device (None: plain round trip)
    def fwd_mul_inv(self, d_a, d_m, batch, stream=None):
        _check(lib.ntt_b200_fwd_mul_inv_batch(self._h, _ptr(d_a), _ptr(d_m), batch, _stream_ptr(stream)),
               "fwd_mul_inv_batch")

    def fwd_mul_inv_host(self, h_a, d_m, batch):
        _check(lib.ntt_b200_fwd_mul_inv_batch_host(self._h, _ptr(h_a), _ptr(d_m), batch), "fwd_mul_inv_batch_host")


class MultiPlan:
    """One plan per device of a device list (ntt_b200_multi_*): the batch is sharded contiguously, no collective."""

    def __init__(self, N, q, psi, devices):
        devs = (C.c_int * len(devices))(*devices)
        h = C.c_void_p()
        if lib.ntt_b200_multi_create(C.byref(h), devs, len(devices), N, q, psi) != 0:
            raise NttError("multi_create: %s" % lib.ntt_b200_multi_last_error().decode())
        self._h, self.N, self.q, self.devices = h, N, q, list(devices)

    def _chk(self, rc, what):
        if rc != 0:
            raise NttError("%s: %s" % (what, lib.ntt_b200_multi_last_error().decode()))

    def close(self):
        if self._h:
            lib.ntt_b200_multi_destroy(self._h)
            self._h = None

    def shard(self, batch, index):
        f, c = C.c_size_t(), C.c_size_t()
        lib.ntt_b200_shard_range(batch, len(self.devices), index, C.byref(f), C.byref(c))
        return int(f.value), int(c.value)

    def _lists(self, d_a, batch, streams):
        n = len(self.devices)
        pa = (C.c_void_p * n)(*[_ptr(x) for x in d_a])
        pb = (C.c_size_t * n)(*batch)
        ps = (C.c_void_p * n)(*[_stream_ptr(s) for s in (streams or [None] * n)])
        return pa, pb, ps

    def fwd(self, d_a, batch, streams=None):
        self._chk(lib.ntt_b200_multi_fwd_batch(self._h, *self._lists(d_a, batch, streams)), "multi_fwd_batch")

    def inv(self, d_a, batch, streams=None):
        self._chk(lib.ntt_b200_multi_inv_batch(self._h, *self._lists(d_a, batch, streams)), "multi_inv_batch")

    def sync(self):
        self._chk(lib.ntt_b200_multi_sync(self._h), "multi_sync")

    def fwd_host(self, h_a, batch):
        self._chk(lib.ntt_b200_multi_fwd_batch_host(self._h, _ptr(h_a), batch), "multi_fwd_batch_host")

    def inv_host(self, h_a, batch):
        self._chk(lib.ntt_b200_multi_inv_batch_host(self._h, _ptr(h_a), batch), "multi_inv_batch_host")

    def fwd_mul_inv_host(self, h_a, d_m, batch):
        pm = None if d_m is None else (C.c_void_p * len(self.devices))(*[_ptr(x) for x in d_m])
        self._chk(lib.ntt_b200_multi_fwd_mul_inv_batch_host(self._h, _ptr(h_a), pm, batch), "multi_fwd_mul_inv_host")


def rns_multi(plans, d_limbs, batch_per_limb, inverse=False):
    """Limbs sharded over devices: plans[l] on any device, d_limbs[l] that limb's polynomials on the same device."""
    n = len(plans)
    arr = (C.c_void_p * n)(*[p._h for p in plans])
    ptr = (C.c_void_p * n)(*[_ptr(x) for x in d_limbs])
    fn = lib.ntt_b200_inv_rns_multi if inverse else lib.ntt_b200_fwd_rns_multi
    if fn(arr, n, ptr, batch_per_limb) != 0:
        raise NttError("rns_multi: %s" % lib.ntt_b200_multi_last_error().decode())


def fwd_rns(plans, d_a, batch_per_limb, stream=None):
    arr = (C.c_void_p * len(plans))(*[p._h for p in plans])
    _check(lib.ntt_b200_fwd_rns(arr, len(plans), _ptr(d_a), batch_per_limb, _stream_ptr(stream)), "fwd_rns")


def inv_rns(plans, d_a, batch_per_limb, stream=None):
    arr = (C.c_void_p * len(plans))(*[p._h for p in plans])
    _check(lib.ntt_b200_inv_rns(arr, len(plans), _ptr(d_a), batch_per_limb, _stream_ptr(stream)), "inv_rns")


# ---- reference-shaped entry points (host numpy arrays, in place) -----------------------------------

def fwd_ntt_ref_harvey(a, N, q, w, w_con):
    _check(lib.ntt_b200_fwd_ntt_ref_harvey(a, N, q, w, w_con), "fwd_ntt_ref_harvey")


def fwd_ntt_ref_harvey_lazy(a, N, q, w, w_con):
    _check(lib.ntt_b200_fwd_ntt_ref_harvey_lazy(a, N, q, w, w_con), "fwd_ntt_ref_harvey_lazy")


def fwd_ntt_ref_harvey_dbl(a1, a2, N, q, w, w_con):
    _check(lib.ntt_b200_fwd_ntt_ref_harvey_dbl(a1, a2, N, q, w, w_con), "fwd_ntt_ref_harvey_dbl")


def inv_ntt_ref_harvey(a, N, q, n_inv, n_inv_con, word_size, w, w_con):
    _check(lib.ntt_b200_inv_ntt_ref_harvey(a, N, q, n_inv, n_inv_con, word_size, w, w_con), "inv_ntt_ref_harvey")


def dropin_reset():
    lib.ntt_b200_dropin_reset()
