"""Kernel time of a given build of libntt_b200.so (plain ctypes, host clock around 20 launches + sync).
usage: python tools/time_lib.py <lib> [logn] [check]   -- for A/B experiments with scratch builds under build/"""
import ctypes as C, os, sys, time, numpy as np
L = C.CDLL(sys.argv[1])
logn = int(sys.argv[2]) if len(sys.argv) > 2 else 14
u64, vp = C.c_uint64, C.c_void_p
L.ntt_b200_plan_create_psi.argtypes = [C.POINTER(vp), C.c_int, u64, u64, u64]
L.ntt_b200_device_alloc.argtypes = [C.c_int, C.POINTER(vp), C.c_size_t]
L.ntt_b200_memcpy_h2d.argtypes = [C.c_int, vp, vp, C.c_size_t]
L.ntt_b200_memcpy_d2h.argtypes = [C.c_int, vp, vp, C.c_size_t]
L.ntt_b200_fwd_batch.argtypes = [vp, vp, C.c_size_t, vp]
L.ntt_b200_inv_batch.argtypes = [vp, vp, C.c_size_t, vp]
L.ntt_b200_device_sync.argtypes = [C.c_int]
N, q = 1 << logn, 0x1FFFFFC800001
x = next(x for x in range(2, 100) if pow(x, (q - 1) // 2, q) == q - 1)
psi = pow(x, (q - 1) // (2 * N), q)  # a primitive 2N-th root
assert pow(psi, N, q) == q - 1
batch = (1 << 26) >> logn
plan = vp(); assert L.ntt_b200_plan_create_psi(C.byref(plan), 0, N, q, psi) == 0
a = np.random.default_rng(1).integers(0, q, size=(batch, N), dtype=np.uint64)
d = vp(); L.ntt_b200_device_alloc(0, C.byref(d), a.nbytes)
L.ntt_b200_memcpy_h2d(0, d, a.ctypes.data, a.nbytes)
out = []
for name, fn in (("fwd", L.ntt_b200_fwd_batch), ("inv", L.ntt_b200_inv_batch)):
    for _ in range(5): fn(plan, d, batch, None)
    L.ntt_b200_device_sync(0); t0 = time.perf_counter()
    for _ in range(20): fn(plan, d, batch, None)
    L.ntt_b200_device_sync(0); out.append("%s %.4f ms" % (name, (time.perf_counter() - t0) / 20 * 1e3))
print(sys.argv[1], "N=2^%d batch %d:" % (logn, batch), "  ".join(out))
if len(sys.argv) > 3:
    L.ntt_b200_memcpy_h2d(0, d, a.ctypes.data, a.nbytes)
    L.ntt_b200_fwd_batch(plan, d, batch, None); L.ntt_b200_inv_batch(plan, d, batch, None)
    r = np.empty_like(a); L.ntt_b200_memcpy_d2h(0, r.ctypes.data, d, a.nbytes)
    print("  round trip mismatches:", int((r != a).sum()))
