"""Element-partition parallelism: one process per GPU, shared-dof exchange over NCCL.

Mirrors what the reference gets from MFEM's ParMesh / ParFiniteElementSpace (ex4.cpp:85,99-101,136):
every interface dof has one OWNER rank (the lowest rank that holds it);
    reduce_to_owner(y)       = P^T : sharers -> owner, summed in ascending rank order (deterministic)
    broadcast_from_owner(x)  = P   : owner -> sharers
Messages are neighbour point-to-point (grouped isend/irecv = ncclSend/ncclRecv inside one
ncclGroup), packed/unpacked by the CUDA kernels madb_pack / madb_unpack on the context stream.
On CPU tensors (gloo, used by the CPU tests of this host logic) packing is plain torch indexing.
"""
import numpy as np
import torch
import torch.distributed as dist


def rank_grid(world):
    """Block layout of the weak-scaled runs: 1x1 / 2x1 / 2x2 / 4x2."""
    px = {1: 1, 2: 2, 4: 2, 8: 4}.get(world)
    if px is None:
        px = int(np.floor(np.sqrt(world)))
        while world % px:
            px -= 1
    return px, world // px


def cartesian_block(rank, world, n, p, lengths=(1.0, 1.0)):
    """This rank's n x n block of the (px*n) x (py*n) mesh and the local->global dof map of the
    order-p H1 space (global lexicographic numbering on the fine node grid)."""
    from . import meshgen as G
    px, py = rank_grid(world)
    rx, ry = rank % px, rank // px
    mesh = G.cartesian_mesh((n, n), lengths=lengths)
    mesh["coords"] = mesh["coords"] + np.array([rx * lengths[0], ry * lengths[1]])
    space = G.h1_space(mesh, p)
    ng = n * p + 1
    NG = px * n * p + 1
    iy, ix = np.divmod(np.arange(space["ndofs"], dtype=np.int64), ng)
    l2g = (iy + ry * n * p) * NG + (ix + rx * n * p)
    boundary = (ix == 0) | (ix == ng - 1) | (iy == 0) | (iy == ng - 1)
    return dict(mesh=mesh, space=space, l2g=l2g, candidates=np.nonzero(boundary)[0], px=px, py=py, rx=rx, ry=ry)


class SharedDofExchange:
    """candidates: local dof indices that may be shared; l2g: their global ids come from l2g[candidates]."""

    def __init__(self, l2g, candidates, device, ctx=None, group=None):
        self.group, self.device, self.ctx = group, torch.device(device), ctx
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        cand = np.asarray(candidates, dtype=np.int64)
        gids = np.asarray(l2g, dtype=np.int64)[cand]
        order = np.argsort(gids)
        gids, cand = gids[order], cand[order]
        allg = [None] * self.world
        dist.all_gather_object(allg, gids, group=group)
        # sharers of every candidate dof
        self.nbr = {}  # peer -> local indices (ascending global id: both sides agree on the order)
        owner = np.full(gids.size, self.rank, dtype=np.int64)
        for r in range(self.world):
            if r == self.rank:
                continue
            common, ia, _ = np.intersect1d(gids, allg[r], assume_unique=True, return_indices=True)
            if common.size:
                self.nbr[r] = cand[ia]
                owner[ia] = np.minimum(owner[ia], r)
        self.owner_of_candidate = dict(zip(cand.tolist(), owner.tolist()))
        own = {c: o for c, o in zip(cand, owner)}
        # P^T: I send dofs I do not own to their owner; I receive (as owner) from every sharer
        self.send_up, self.recv_up = {}, {}
        for r, idx in self.nbr.items():
            mine = np.array([own[i] == self.rank for i in idx], dtype=bool)
            theirs = np.array([own[i] == r for i in idx], dtype=bool)
            if theirs.any():
                self.send_up[r] = idx[theirs]
            if mine.any():
                self.recv_up[r] = idx[mine]
        self.peers = sorted(self.nbr)
        self._bufs = {}
        self.n_owned_shared = int(sum(1 for c in cand if own[c] == self.rank and any(c in v for v in self.nbr.values())))
        self.is_owner_mask = None

    def owned_mask(self, ndofs):
        """True for dofs whose owner is this rank (true dofs)."""
        m = np.ones(ndofs, dtype=bool)
        for c, o in self.owner_of_candidate.items():
            if o != self.rank:
                m[c] = False
        return m

    # ---- packing -------------------------------------------------------------------------
    def _idx(self, key, arr):
        k = (key, id(arr))
        if k not in self._bufs:
            self._bufs[k] = torch.from_numpy(np.ascontiguousarray(arr, dtype=np.int32)).to(self.device)
        return self._bufs[k]

    def _buf(self, key, n):
        k = ("buf", key)
        if k not in self._bufs or self._bufs[k].numel() != n:
            self._bufs[k] = torch.empty(n, dtype=torch.float64, device=self.device)
        return self._bufs[k]

    def _pack(self, vec, idx_t, out):
        if vec.is_cuda:
            from . import lib, _check
            _check(lib().madb_pack(self.ctx.h, idx_t.numel(), idx_t.data_ptr(), vec.data_ptr(), out.data_ptr()))
        else:
            torch.index_select(vec, 0, idx_t.long(), out=out)

    def _unpack(self, vec, idx_t, src, add):
        if vec.is_cuda:
            from . import lib, _check
            _check(lib().madb_unpack(self.ctx.h, idx_t.numel(), idx_t.data_ptr(), src.data_ptr(), vec.data_ptr(), int(add)))
        elif add:
            vec.index_add_(0, idx_t.long(), src)
        else:
            vec.index_copy_(0, idx_t.long(), src)

    def _exchange(self, vec, send, recv, add, tag):
        """One all-to-all-v (ncclGroup of send/recv pairs in which EVERY rank takes part, empty messages for
        non-neighbours) instead of hand-built point-to-point batches: the same traffic, but the collective is
        symmetric across ranks, which NCCL's lazy connection set-up requires at more than two ranks.
        Pack and unpack are one kernel each; a dof received from several sharers (a corner) is one destination
        with its sources in ascending peer rank: fixed summation order."""
        key = ("a2a", tag)
        if key not in self._bufs:
            scount = [len(send[r]) if r in send else 0 for r in range(self.world)]
            rcount = [len(recv[r]) if r in recv else 0 for r in range(self.world)]
            sidx = np.concatenate([send[r] for r in range(self.world) if r in send]) if sum(scount) else np.zeros(0, np.int64)
            ridx = np.concatenate([recv[r] for r in range(self.world) if r in recv]) if sum(rcount) else np.zeros(0, np.int64)
            # destinations with their receive-buffer positions (ascending position = ascending peer)
            order = np.argsort(ridx, kind="stable")
            dst, start = np.unique(ridx[order], return_index=True)
            counts = np.diff(np.append(start, ridx.size))
            assert counts.size == 0 or counts.max() <= 4, "a dof shared by more than 5 ranks"
            src4 = np.full((dst.size, 4), -1, dtype=np.int32)
            for k in range(4):
                m = counts > k
                src4[m, k] = order[start[m] + k]
            self._bufs[key] = (scount, rcount,
                               torch.from_numpy(np.ascontiguousarray(sidx, dtype=np.int32)).to(self.device),
                               torch.from_numpy(np.ascontiguousarray(dst, dtype=np.int32)).to(self.device),
                               torch.from_numpy(src4).to(self.device),
                               torch.empty(max(sum(scount), 1), dtype=torch.float64, device=self.device),
                               torch.empty(max(sum(rcount), 1), dtype=torch.float64, device=self.device))
        scount, rcount, sidx, dst, src4, sbuf, rbuf = self._bufs[key]
        ns, nr = sum(scount), sum(rcount)
        if vec.is_cuda and self.ctx is not None and torch.cuda.current_stream(self.device).cuda_stream != self.ctx.stream:
            # pack / unpack run on the context's (non-blocking) stream, NCCL on torch's current stream: make them the
            # same stream for the duration of the exchange, ordered after the work already queued by the caller
            ext = torch.cuda.ExternalStream(self.ctx.stream, device=self.device)
            ext.wait_stream(torch.cuda.current_stream(self.device))
            with torch.cuda.stream(ext):
                self._exchange(vec, send, recv, add, tag)
            torch.cuda.current_stream(self.device).wait_stream(ext)
            return
        if ns:
            self._pack(vec, sidx, sbuf[:ns])
        dist.all_to_all_single(rbuf[:nr], sbuf[:ns], rcount, scount, group=self.group)
        if nr:
            if vec.is_cuda:
                from . import lib, _check
                _check(lib().madb_unpack_multi(self.ctx.h, dst.numel(), src4.data_ptr(), dst.data_ptr(), rbuf.data_ptr(),
                                               vec.data_ptr(), int(add)))
            else:
                d = dst.long()
                acc = vec[d].clone() if add else torch.zeros(d.numel(), dtype=vec.dtype)
                for k in range(4):
                    m = src4[:, k] >= 0
                    acc[m] = acc[m] + rbuf[src4[m, k].long()]
                vec[d] = acc

    def reduce_to_owner(self, y):
        """P^T: the owner's copy becomes the sum over all sharers (own value first, then ascending rank)."""
        self._exchange(y, self.send_up, self.recv_up, True, "up")

    def broadcast_from_owner(self, x):
        """P: every sharer's copy is overwritten with the owner's value."""
        self._exchange(x, self.recv_up, self.send_up, False, "down")


# ---------------------------------------------------------------------------------------------------------------
# Overlapping element partition: owned elements + one layer of ghost elements.
# Every element that touches a dof owned by this rank is local, so the ROWS of owned dofs -- residual entries and
# Jacobian rows, i.e. this rank's rows of P^T A P -- are complete after the local assembly and summed in the local,
# fixed patch order; the only communication per assembly is P: owner -> copies for the state and the parameter
# fields (HaloExchange.forward).  P^T (reverse) is available for vectors assembled on non-overlapping partitions.
# ---------------------------------------------------------------------------------------------------------------
def halo_lists(l2g, owner, rank, world, group=None):
    """Neighbour lists of the exchange from the local->global map and the owner rank of every local dof:
    own[r]  = local indices of dofs this rank owns and rank r holds a copy of,
    ghost[r] = local indices of the copies this rank holds of dofs rank r owns; both in ascending global id."""
    l2g = np.asarray(l2g, dtype=np.int64)
    owner = np.asarray(owner, dtype=np.int64)
    need = {}
    for r in np.unique(owner):
        r = int(r)
        if r == rank:
            continue
        idx = np.nonzero(owner == r)[0]
        idx = idx[np.argsort(l2g[idx], kind="stable")]
        need[r] = idx
    published = [None] * world
    dist.all_gather_object(published, {r: l2g[idx] for r, idx in need.items()}, group=group)
    order = np.argsort(l2g, kind="stable")
    sorted_g = l2g[order]
    own = {}
    for r in range(world):
        if r == rank or published[r] is None or rank not in published[r]:
            continue
        g = published[r][rank]
        pos = np.searchsorted(sorted_g, g)
        if np.any(pos >= sorted_g.size) or np.any(sorted_g[np.minimum(pos, sorted_g.size - 1)] != g):
            raise ValueError("halo_lists: rank %d asks rank %d for dofs it does not hold" % (r, rank))
        loc = order[pos]
        if np.any(owner[loc] != rank):
            raise ValueError("halo_lists: rank %d asks rank %d for dofs it does not own" % (r, rank))
        own[r] = loc
    return own, need


def lowest_rank_owner(l2g, candidates, ndofs, rank, world, group=None):
    """Owner rank of every local dof of a NON-overlapping partition: the lowest rank that holds it (candidates = local
    indices of the dofs that may be shared, e.g. those on the block boundary)."""
    l2g = np.asarray(l2g, dtype=np.int64)
    cand = np.asarray(candidates, dtype=np.int64)
    allg = [None] * world
    dist.all_gather_object(allg, l2g[cand], group=group)
    owner = np.full(ndofs, rank, dtype=np.int64)
    for r in range(rank):
        hit = cand[np.isin(l2g[cand], allg[r])]
        owner[hit] = np.minimum(owner[hit], r)
    return owner


class HaloExchange:
    """P (forward: owner -> copies) and P^T (reverse: copies -> owner, added in ascending peer rank) for one local
    vector.  CUDA tensors go through the C ABI (madb_exchange_*: pack kernel, ncclSend/ncclRecv group on a
    communication stream, unpack kernel; asynchronous on the context stream); CPU tensors (gloo, the tests of the host
    logic) through torch.distributed."""

    def __init__(self, own, ghost, ctx=None, comm=None, group=None):
        self.group, self.ctx, self.comm = group, ctx, comm
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        self.own = {int(r): np.asarray(v, dtype=np.int32) for r, v in sorted(own.items()) if len(v)}
        self.ghost = {int(r): np.asarray(v, dtype=np.int32) for r, v in sorted(ghost.items()) if len(v)}
        self.h = None
        if comm is not None:
            import ctypes as C
            from . import lib, _check

            def pack(d):
                peers = np.array(sorted(d), dtype=np.int32)
                counts = np.array([d[r].size for r in peers], dtype=np.int32)
                idx = np.concatenate([d[r] for r in peers]).astype(np.int32) if peers.size else np.zeros(0, np.int32)
                return peers, counts, np.ascontiguousarray(idx)
            op, oc, oi = pack(self.own)
            gp, gc, gi = pack(self.ghost)
            h = C.c_void_p()
            _check(lib().madb_exchange_create(comm.h, op.size, op.ctypes.data, oc.ctypes.data, oi.ctypes.data,
                                              gp.size, gp.ctypes.data, gc.ctypes.data, gi.ctypes.data, C.byref(h)))
            self.h = h

    # ---- device path ----
    def begin(self, vec, reverse=False):
        from . import lib, _check
        _check(lib().madb_exchange_begin(self.h, vec.data_ptr(), 1 if reverse else 0))

    def end(self, vec, add=False):
        from . import lib, _check
        _check(lib().madb_exchange_end(self.h, vec.data_ptr(), 1 if add else 0))

    # ---- both paths ----
    def forward(self, x):
        """P: the copies are overwritten with the owner's values."""
        if x.is_cuda:
            self.begin(x, False)
            self.end(x, False)
        else:
            self._cpu(x, self.own, self.ghost, False)

    def reverse(self, y):
        """P^T: the owner's value becomes own + sum of the copies (ascending peer rank)."""
        if y.is_cuda:
            self.begin(y, True)
            self.end(y, True)
        else:
            self._cpu(y, self.ghost, self.own, True)

    def _cpu(self, vec, send, recv, add):
        scount = [send[r].size if r in send else 0 for r in range(self.world)]
        rcount = [recv[r].size if r in recv else 0 for r in range(self.world)]
        sidx = np.concatenate([send[r] for r in sorted(send)]) if send else np.zeros(0, np.int64)
        sbuf = vec[torch.from_numpy(sidx.astype(np.int64))].contiguous() if sum(scount) else torch.zeros(0, dtype=vec.dtype)
        rbuf = torch.zeros(sum(rcount), dtype=vec.dtype)
        dist.all_to_all_single(rbuf, sbuf, rcount, scount, group=self.group)
        off = 0
        for r in sorted(recv):  # ascending peer rank: fixed summation order
            idx = torch.from_numpy(recv[r].astype(np.int64))
            part = rbuf[off:off + recv[r].size]
            if add:
                vec[idx] = vec[idx] + part
            else:
                vec[idx] = part
            off += recv[r].size

    def __del__(self):
        try:
            if self.h is not None:
                from . import lib
                lib().madb_exchange_destroy(self.h)
        except Exception:
            pass


class Comm:
    """NCCL communicator behind the C ABI (madb_comm_*); the unique id travels through torch.distributed here, a C++ host
    would use MPI_Bcast."""

    def __init__(self, ctx, group=None):
        import ctypes as C
        from . import lib, _check
        rank, world = dist.get_rank(group), dist.get_world_size(group)
        buf = C.create_string_buffer(128)
        if rank == 0:
            _check(lib().madb_comm_unique_id(buf))
        box = [bytes(buf.raw)]
        dist.broadcast_object_list(box, src=0, group=group)
        h = C.c_void_p()
        _check(lib().madb_comm_create(ctx.h, box[0], rank, world, C.byref(h)))
        self.h, self.ctx, self.rank, self.world = h, ctx, rank, world

    def allreduce_sum(self, values):
        from . import lib, _check
        v = np.ascontiguousarray(values, dtype=np.float64)
        _check(lib().madb_comm_allreduce_sum(self.h, v.size, v.ctypes.data))
        return v

    def __del__(self):
        try:
            from . import lib
            lib().madb_comm_destroy(self.h)
        except Exception:
            pass


def cartesian_block_ghost(rank, world, n, p_h1, p_l2=None, lengths=(1.0, 1.0)):
    """This rank's n x n block of the (px*n) x (py*n) mesh plus one layer of ghost elements towards the higher ranks
    (right / top), the order-p_h1 H1 space and (optionally) the order-p_l2 L2 space on it.

    A dof is owned by the lowest rank whose OWN elements touch it: block (rx, ry) owns the dofs of its closed block
    except those on its left / bottom edge when a lower neighbour exists; the ghost layer makes every element touching
    an owned dof local.  Returns dict(mesh, h1, l2, l2g_h1, owner_h1, l2g_l2, owner_l2, owned_elements, px, py)."""
    from . import meshgen as G
    px, py = rank_grid(world)
    rx, ry = rank % px, rank // px
    gx, gy = (1 if rx < px - 1 else 0), (1 if ry < py - 1 else 0)
    nx, ny = n + gx, n + gy
    mesh = G.cartesian_mesh((nx, ny), lengths=(lengths[0] * nx / n, lengths[1] * ny / n))
    mesh["coords"] = mesh["coords"] + np.array([rx * lengths[0], ry * lengths[1]])
    h1 = G.h1_space(mesh, p_h1)
    p = p_h1
    ngx = nx * p + 1
    NGX = px * n * p + 1
    iy, ix = np.divmod(np.arange(h1["ndofs"], dtype=np.int64), ngx)
    gix, giy = ix + rx * n * p, iy + ry * n * p
    l2g_h1 = giy * NGX + gix

    def block_owner(gi_, nb, last):  # block index of the lowest block touching global node index gi_ (block size nb*p)
        b = gi_ // (n * p)
        on_edge = (gi_ % (n * p) == 0) & (b > 0)
        b = np.where(on_edge, b - 1, b)
        return np.minimum(b, last)
    owner_h1 = block_owner(giy, n, py - 1) * px + block_owner(gix, n, px - 1)
    ey, ex = np.divmod(np.arange(nx * ny, dtype=np.int64), nx)
    owned_el = (ex < n) & (ey < n)
    owner_el = (np.minimum((ey + ry * n) // n, py - 1)) * px + np.minimum((ex + rx * n) // n, px - 1)
    out = dict(mesh=mesh, h1=h1, l2g_h1=l2g_h1, owner_h1=owner_h1, owned_elements=owned_el, px=px, py=py, rx=rx, ry=ry,
               owner_el=owner_el)
    if p_l2 is not None:
        l2 = G.l2_space(mesh, p_l2)
        nd = (p_l2 + 1) ** 2
        gel = (ey + ry * n) * (px * n) + (ex + rx * n)
        out["l2"] = l2
        out["l2g_l2"] = (gel[:, None] * nd + np.arange(nd)[None, :]).reshape(-1)
        out["owner_l2"] = np.repeat(owner_el, nd)
    return out
