"""Per-phase cycle counts of the FP64 ring kernel (needs a library built with -DNTT_RING_TRACE, path in argv[1]).
usage: python tools/trace_phases.py <libntt_b200.so built with -DNTT_RING_TRACE> fwd|inv"""
import ctypes as C, sys, numpy as np
L = C.CDLL(sys.argv[1])
u64, vp = C.c_uint64, C.c_void_p
L.ntt_b200_plan_create_psi.argtypes = [C.POINTER(vp), C.c_int, u64, u64, u64]
L.ntt_b200_device_alloc.argtypes = [C.c_int, C.POINTER(vp), C.c_size_t]
L.ntt_b200_memcpy_h2d.argtypes = [C.c_int, vp, vp, C.c_size_t]
L.ntt_b200_fwd_batch.argtypes = [vp, vp, C.c_size_t, vp]
L.ntt_b200_inv_batch.argtypes = [vp, vp, C.c_size_t, vp]
L.ntt_cuda_trace_read.argtypes = [vp, C.c_size_t]
N, q, psi, batch = 1 << 14, 0x1FFFFFC800001, 20456969886, 4096
plan = vp(); assert L.ntt_b200_plan_create_psi(C.byref(plan), 0, N, q, psi) == 0
a = np.random.default_rng(1).integers(0, q, size=(batch, N), dtype=np.uint64)
d = vp(); L.ntt_b200_device_alloc(0, C.byref(d), a.nbytes)
L.ntt_b200_memcpy_h2d(0, d, a.ctypes.data, a.nbytes)
which = sys.argv[2] if len(sys.argv) > 2 else "fwd"
fn = L.ntt_b200_fwd_batch if which == "fwd" else L.ntt_b200_inv_batch
for _ in range(3): fn(plan, d, batch, None)
t = np.zeros(16 * 64 * 8, dtype=np.int64); L.ntt_cuda_trace_read(t.ctypes.data, t.size)
t = t.reshape(16, 64, 8)
names = {"fwd": ["wait0", "wait1", "A_end", "sync_end", "B_end", "C1_end", "C2_end", "rearm_end"],
         "inv": ["wait0", "wait_lo", "C1_end", "C2_end", "B_end", "sync_end", "A_loaded", "A_end"]}[which]
for w in (0, 7, 15):
    print("warp", w)
    for k in range(8, 12):
        ev = t[w, k]; base = t[0, k, 0]
        print("  poly %2d start %+6d |" % (k, ev[0] - base), " ".join("%s %5d" % (names[i], ev[i] - ev[i - 1]) for i in range(1, 8)), "| total", ev[7] - ev[0])
print("cycles per polynomial (warp 0):", (t[0, 20, 0] - t[0, 8, 0]) / 12.0)
