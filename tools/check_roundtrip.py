"""forward+inverse round trip of 256 polynomials through a given build of libntt_b200.so (plain ctypes)."""
import ctypes as C, os, sys, numpy as np
L = C.CDLL(sys.argv[1])
u64, vp = C.c_uint64, C.c_void_p
L.ntt_b200_plan_create_psi.argtypes = [C.POINTER(vp), C.c_int, u64, u64, u64]
L.ntt_b200_device_alloc.argtypes = [C.c_int, C.POINTER(vp), C.c_size_t]
L.ntt_b200_memcpy_h2d.argtypes = [C.c_int, vp, vp, C.c_size_t]
L.ntt_b200_memcpy_d2h.argtypes = [C.c_int, vp, vp, C.c_size_t]
L.ntt_b200_fwd_batch.argtypes = [vp, vp, C.c_size_t, vp]
L.ntt_b200_inv_batch.argtypes = [vp, vp, C.c_size_t, vp]
N, q, psi, batch = 1 << 14, 0x1FFFFFC800001, 20456969886, 256
plan = vp(); assert L.ntt_b200_plan_create_psi(C.byref(plan), 0, N, q, psi) == 0
a = np.random.default_rng(1).integers(0, q, size=(batch, N), dtype=np.uint64)
d = vp(); L.ntt_b200_device_alloc(0, C.byref(d), a.nbytes)
L.ntt_b200_memcpy_h2d(0, d, a.ctypes.data, a.nbytes)
L.ntt_b200_fwd_batch(plan, d, batch, None); L.ntt_b200_inv_batch(plan, d, batch, None)
r = np.empty_like(a); L.ntt_b200_memcpy_d2h(0, r.ctypes.data, d, a.nbytes)
bad = np.argwhere(r != a)
print(sys.argv[1], "mismatches", len(bad), "rows", sorted(set(bad[:, 0].tolist()))[:8])
