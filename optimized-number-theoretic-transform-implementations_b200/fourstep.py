"""One negacyclic NTT of size N spread over G = 2^g GPUs (BASELINE config 5: N = 2^22, one exchange step).

Decomposition (restating the loop nest of src/ntt_reference.c:19-30): with the
coefficients held CYCLICALLY -- rank p owns a[p + G*k] -- every stage whose butterfly distance is >= G touches
one residue class only, and on rank p's slice those log2(N/G) stages are exactly a complete forward NTT of size
N/G with root psi^G (same twiddles for every p).  After ONE all-to-all that re-lays the array out in CONTIGUOUS
blocks -- rank r owns a[r*N/G .. (r+1)*N/G) -- the remaining g stages (distances G/2 .. 1) are local again and
finish the transform: rank r then holds words [r*N/G, (r+1)*N/G) of fwd_ntt_ref_harvey's output (bit-reversed
order, fully reduced).  The inverse is the mirror image: tail stages on contiguous blocks, all-to-all back to
cyclic slices, complete size-N/G inverse whose scale is the global N^-1.

The local transforms run on the same sm_100a kernels as the batched path (C-ABI: ntt_b200_fwd_batch,
ntt_b200_fwd_tail_block, ...); torch.distributed (NCCL over NVLink) provides the all-to-all.  One process per GPU.
"""
import importlib

import numpy as np

_pkg = importlib.import_module(__name__.rsplit(".", 1)[0])


def exchange_cyclic_to_blocks(local, world, dist=None, group=None):
    """local: this rank's cyclic slice (N/G words, 1-D torch int64 tensor).  Returns this rank's contiguous block.

    Rank p's slice index k holds position e = p + G*k.  Block r needs k in [r*N/G^2, (r+1)*N/G^2) from every p,
    and position e sits at block offset p + G*(k - r*N/G^2): the received [G, N/G^2] pieces are interleaved.
    """
    import torch
    G = world
    n_local = local.numel()
    assert n_local % G == 0
    piece = n_local // G
    if G == 1:
        return local.clone()
    recv = torch.empty_like(local)
    if dist.get_backend(group) == "nccl":
        dist.all_to_all_single(recv, local.contiguous(), group=group)
    else:  # gloo (CPU tests): all_gather and pick this rank's piece from every source
        rank = dist.get_rank(group)
        gathered = [torch.empty_like(local) for _ in range(G)]
        dist.all_gather(gathered, local.contiguous(), group=group)
        recv = torch.cat([g[rank * piece:(rank + 1) * piece] for g in gathered])
    # recv[p*piece + kk] = a[p + G*(r*piece + kk)]  ->  block[kk*G + p]
    return recv.view(G, piece).t().contiguous().view(-1)


def exchange_blocks_to_cyclic(block, world, dist=None, group=None):
    """Inverse of exchange_cyclic_to_blocks."""
    import torch
    G = world
    n_local = block.numel()
    piece = n_local // G
    if G == 1:
        return block.clone()
    send = block.view(piece, G).t().contiguous().view(-1)  # send[p*piece + kk] = block[kk*G + p] -> rank p
    recv = torch.empty_like(send)
    if dist.get_backend(group) == "nccl":
        dist.all_to_all_single(recv, send, group=group)
    else:
        rank = dist.get_rank(group)
        gathered = [torch.empty_like(send) for _ in range(G)]
        dist.all_gather(gathered, send, group=group)
        recv = torch.cat([g[rank * piece:(rank + 1) * piece] for g in gathered])
    # from source r: kk-th word is slice index r*piece + kk
    return recv


class DistributedNtt:
    """Plans of one rank for a size-N transform over `world` = 2^g ranks."""

    def __init__(self, N, q, psi, rank, world, device=0):
        assert world & (world - 1) == 0 and world >= 1
        self.N, self.q, self.psi, self.rank, self.world = N, q, psi, rank, world
        self.g = world.bit_length() - 1
        assert (N // world) >= world, "need N >= G^2"
        self.n_local = N // world
        # complete transforms of the cyclic slices: size N/G, root psi^G
        self.local = _pkg.Plan.from_psi(self.n_local, q, _pkg.pow_mod(psi, world, q), device=device)
        # the local inverse must scale by the GLOBAL N^-1 (= (N/G)^-1 * G^-1)
        self.local.set_inverse_scale(_pkg.pow_mod((q + 1) // 2, N.bit_length() - 1, q))
        # tail stages use the tables of the full-size transform
        self.full = _pkg.Plan.from_psi(N, q, psi, device=device) if world > 1 else None

    def forward(self, slice_dev, dist=None, group=None, stream=None):
        """slice_dev: this rank's cyclic slice a[rank + G*k] on the GPU (modified).  Returns this rank's block of
        the transform (words [rank*N/G, (rank+1)*N/G) of the reference output)."""
        self.local.fwd(slice_dev, 1, stream)
        if self.world == 1:
            return slice_dev
        block = exchange_cyclic_to_blocks(slice_dev, self.world, dist, group)
        self.full.fwd_tail_block(block, self.g, self.rank, stream)
        return block

    def inverse(self, block_dev, dist=None, group=None, stream=None):
        """block_dev: this rank's contiguous block of the NTT-domain array.  Returns its cyclic slice of the
        coefficient array, scaled by N^-1 and fully reduced (inv_ntt_ref_harvey semantics)."""
        if self.world == 1:
            self.local.inv(block_dev, 1, stream)
            return block_dev
        self.full.inv_tail_block(block_dev, self.g, self.rank, stream)
        sl = exchange_blocks_to_cyclic(block_dev, self.world, dist, group)
        self.local.inv(sl, 1, stream)
        return sl

    def close(self):
        self.local.close()
        if self.full is not None:
            self.full.close()


class PeerExchange:
    """The exchange step without a collective: every rank's slice buffer (cudaMalloc, N/G words) is mapped into
    every other rank's address space through CUDA IPC, so the tail kernels load / store the group members straight
    from / into peer memory over NVLink (ntt_b200_fwd_tail_gather, ntt_b200_inv_tail_scatter), and the ranks are
    ordered by a flag barrier that runs as a kernel on the caller's stream (ntt_b200_peer_barrier).
    `dist` is only used once, to hand the IPC handles round."""

    def __init__(self, n_local, rank, world, device, dist, group=None, batch=1):
        import ctypes as C
        self.rank, self.world, self.device, self.n_local = rank, world, device, n_local * batch
        n_local = n_local * batch                                  # `batch` slices one after the other
        self.slice_ptr = _pkg.device_alloc(device, n_local * 8)
        self.flags_ptr = _pkg.device_alloc(device, 512)            # `world` uint32 flags, timeout word at +256
        _pkg.memcpy_h2d(device, self.flags_ptr, np.zeros(64, dtype=np.uint64))
        _pkg.device_sync(device)
        mine = (_pkg.ipc_export(device, self.slice_ptr), _pkg.ipc_export(device, self.flags_ptr))
        handles = [None] * world
        dist.all_gather_object(handles, mine, group=group)
        self._opened = []
        self.slices = (C.c_void_p * world)()
        self.flags = (C.c_void_p * world)()
        for p in range(world):
            if p == rank:
                self.slices[p], self.flags[p] = self.slice_ptr, self.flags_ptr
            else:
                sp, fp = _pkg.ipc_open(device, handles[p][0]), _pkg.ipc_open(device, handles[p][1])
                self._opened += [sp, fp]
                self.slices[p], self.flags[p] = sp, fp
        self.epoch = 0
        self._dist, self._group = dist, group
        dist.barrier(group=group)                                  # every rank's flags are zeroed and mapped

    def barrier(self, stream=None):
        """All ranks' work enqueued before this call (on their streams) is complete and visible afterwards.
        The epoch is counted on the device, so the call can be captured in a CUDA graph and replayed."""
        _pkg.peer_barrier(self.device, self.flags, self.flags_ptr, self.rank, self.world, 0,
                          self.flags_ptr + 256, stream)

    def timed_out(self):
        w = np.zeros(1, dtype=np.uint64)
        _pkg.memcpy_d2h(self.device, w, self.flags_ptr + 256)
        return bool(w[0] & 0xFFFFFFFF)

    def load_slice(self, host_u64):
        assert host_u64.size == self.n_local
        _pkg.memcpy_h2d(self.device, self.slice_ptr, np.ascontiguousarray(host_u64, dtype=np.uint64))

    def read_slice(self):
        out = np.empty(self.n_local, dtype=np.uint64)
        _pkg.memcpy_d2h(self.device, out, self.slice_ptr)
        return out

    def close(self):
        _pkg.device_sync(self.device)
        self._dist.barrier(group=self._group)                      # nobody is still reading our buffers
        for ptr in self._opened:
            _pkg.ipc_close(self.device, ptr)
        self._dist.barrier(group=self._group)
        _pkg.device_free(self.device, self.slice_ptr)
        _pkg.device_free(self.device, self.flags_ptr)


class FusedDistributedNtt(DistributedNtt):
    """DistributedNtt whose exchange is fused into the tail kernels (peer loads / stores, no NCCL on the data path).

    The rank's cyclic slice lives in `self.px.slice_ptr` (px.load_slice / px.read_slice move it from / to the host).
    forward(block) : slice -> this rank's block of the transform;  inverse(block) : block -> slice.
    Calls must alternate forward / inverse (each barrier then also fences the previous step's peer accesses);
    to repeat the same direction, call self.px.barrier() in between."""

    def __init__(self, N, q, psi, rank, world, device, dist, group=None, batch=1):
        """batch > 1: `batch` transforms share every launch and every barrier (slice buffer = batch slices of N/G
        words one after the other, block buffer = batch blocks): one N = 2^22 exchange is latency-bound."""
        super().__init__(N, q, psi, rank, world, device)
        assert world > 1
        self.device, self.batch = device, batch
        self.px = PeerExchange(self.n_local, rank, world, device, dist, group, batch)

    def forward(self, block_dev, stream=None):
        self.local.fwd(self.px.slice_ptr, self.batch, stream)
        self.px.barrier(stream)
        self.full.fwd_tail_gather(self.px.slices, block_dev, self.g, self.rank, stream, self.batch)
        return block_dev

    def inverse(self, block_dev, stream=None):
        self.full.inv_tail_scatter(self.px.slices, block_dev, self.g, self.rank, stream, self.batch)
        self.px.barrier(stream)
        self.local.inv(self.px.slice_ptr, self.batch, stream)

    def check(self):
        """Synchronise and raise if any peer barrier gave up waiting (k_peer_barrier reports a timeout instead of
        hanging the GPU; the kernels behind it then ran on stale peer data, so the results must not be used)."""
        _pkg.device_sync(self.device)
        if self.px.timed_out():
            raise _pkg.NttError("peer barrier timed out: a rank never arrived; results of this step are invalid")

    def capture_pair(self, block_dev):
        """forward(block) followed by inverse(block) as one CUDA graph (returns the torch.cuda.CUDAGraph): the
        pair is eight short launches, and at N = 2^22 the host's launch path is slower than the kernels."""
        import torch
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph, stream=side, capture_error_mode="relaxed"):
            st = torch.cuda.current_stream()
            self.forward(block_dev, st)
            self.inverse(block_dev, st)
        return graph

    def close(self):
        try:
            self.check()
        finally:
            self.px.close()
            super().close()


def emulate_forward_single_gpu(N, q, psi, a_host, world):
    """All `world` ranks emulated one after the other on ONE GPU (no collective): the same kernels and the same
    index arithmetic as the multi-process path; used by the single-GPU parity test."""
    import torch
    G = world
    parts = [DistributedNtt(N, q, psi, r, G) for r in range(G)]
    slices = [torch.from_numpy(np.ascontiguousarray(a_host[p::G]).view(np.int64)).cuda() for p in range(G)]
    for p in range(G):
        parts[p].local.fwd(slices[p], 1)
    piece = (N // G) // G
    out = []
    for r in range(G):
        recv = torch.cat([slices[p][r * piece:(r + 1) * piece] for p in range(G)])
        block = recv.view(G, piece).t().contiguous().view(-1)
        if G > 1:
            parts[r].full.fwd_tail_block(block, parts[r].g, r)
        out.append(block)
    res = torch.cat(out).cpu().numpy().view(np.uint64)
    # and back again
    blocks = [o.clone() for o in out]
    for r in range(G):
        if G > 1:
            parts[r].full.inv_tail_block(blocks[r], parts[r].g, r)
    sends = [b.view(piece, G).t().contiguous().view(-1) for b in blocks]
    back = np.empty(N, dtype=np.uint64)
    for p in range(G):
        sl = torch.cat([sends[r][p * piece:(p + 1) * piece] for r in range(G)])
        parts[p].local.inv(sl, 1)
        back[p::G] = sl.cpu().numpy().view(np.uint64)
    for d in parts:
        d.close()
    return res, back
