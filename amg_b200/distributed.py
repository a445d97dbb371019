"""Multi-GPU driver of the solve phase: one process per GPU, torch.distributed for the plumbing.

What shards (DESIGN.md "Multi-GPU"): under the reference's row-order Gauss-Seidel only a level whose
F pass and C pass are each ONE wavefront (level 0 of the 5-/7-point problems: an exact red/black split)
can be partitioned without a cross-device dependency chain.  Every rank holds the resident hierarchy
(HBM is not the constraint: 256^3 needs ~9 of 180 GB); on level 0 each rank runs the Gauss-Seidel
passes, the residual and the prolongation on its own row block and refreshes the ghost x entries it
reads from their owners (point-to-point halo exchange, no collective on the data path except the
residual gather to rank 0 and the broadcast of the coarse correction).  Levels >= 1 are a single
dependency chain per sweep and run on rank 0 ("agglomerated").  The arithmetic of every row is the same
as on one GPU, so the iterates are bit-identical to the single-GPU (and the reference's CPU) iterates.

The driver is backend-agnostic: GpuBackend calls libamgb200.so; tests/test_sharded_gloo.py plugs in a CPU
emulation built on the oracle to check partitioning, ghost lists and the exchange schedule with gloo.
"""
import numpy as np


def split_even(nitems, parts):
    """parts contiguous item ranges of (nearly) equal size"""
    cuts = [(nitems * p) // parts for p in range(parts + 1)]
    return [(cuts[p], cuts[p + 1]) for p in range(parts)]


class Partition:
    """Ownership of level-0 rows in schedule numbering: rank p owns the F rows of F items
    [fi[p], fi[p+1]) and the C rows of C items [ci[p], ci[p+1]) (an item = rows_per_item rows)."""

    def __init__(self, n, nF, itemsF, itemsC, rows_per_item, world):
        self.n, self.nF, self.rpi, self.world = n, nF, rows_per_item, world
        self.f_items = split_even(itemsF, world)
        self.c_items = split_even(itemsC, world)
        # schedule-row ranges per rank and pass
        self.f_rows = [(min(a * rows_per_item, nF), min(b * rows_per_item, nF)) for a, b in self.f_items]
        self.c_rows = [(min(nF + a * rows_per_item, n), min(nF + b * rows_per_item, n)) for a, b in self.c_items]
        self.f_bounds = np.array([r[0] for r in self.f_rows] + [nF])
        self.c_bounds = np.array([r[0] for r in self.c_rows] + [n])

    def owner(self, k):
        """owning rank of schedule rows k (array)"""
        k = np.asarray(k)
        own = np.empty(k.shape, np.int64)
        isF = k < self.nF
        own[isF] = np.searchsorted(self.f_bounds, k[isF], side="right") - 1
        own[~isF] = np.searchsorted(self.c_bounds, k[~isF], side="right") - 1
        return np.clip(own, 0, self.world - 1)

    def rows(self, rank, which):
        a, b = (self.f_rows if which == 0 else self.c_rows)[rank]
        return np.arange(a, b)


def expand_ranges(starts, ends):
    """concatenation of arange(s, e) for all (s, e), vectorised (level 0 of 256^3 has 2 million rows per rank)"""
    lens = ends - starts
    total = int(lens.sum())
    if total == 0:
        return np.zeros(0, np.int64)
    out = np.ones(total, np.int64)
    first = np.cumsum(lens) - lens                      # offset of every non-empty range in the output
    nz = lens > 0
    out[first[nz]] = starts[nz] - np.concatenate(([0], (ends[nz] - 1)[:-1]))
    return np.cumsum(out)


def ghost_lists(A_nat, order, part, rank):
    """For `rank`: ghosts[p][src] = schedule indices of pass-p x entries owned by `src` that this rank's
    rows of the OTHER pass (and, for the residual, of both passes) read.  Symmetric send lists follow by
    evaluating the same function for the peer, which every rank does locally (the matrix is replicated)."""
    n = part.n
    pos = np.empty(n, np.int64)
    pos[order] = np.arange(n)
    out = {}
    for which in (0, 1):                      # ghost entries belonging to pass `which`
        reader_rows = np.concatenate([part.rows(rank, 0), part.rows(rank, 1)])
        nat = order[reader_rows]
        starts, ends = A_nat.row_ptr[nat].astype(np.int64), A_nat.row_ptr[nat + 1].astype(np.int64)
        cols = np.unique(pos[A_nat.col_idx[expand_ranges(starts, ends)]])
        cols = cols[(cols < part.nF) if which == 0 else (cols >= part.nF)]
        own = part.owner(cols)
        out[which] = {src: cols[own == src] for src in range(part.world) if src != rank and (own == src).any()}
    return out


def all_ghost_lists(A_nat, order, part):
    """ghost_lists for every rank in ONE vectorised pass over the matrix: out[rank][which][src] (same content as ghost_lists(.., rank)).
    Every entry (row, col) whose row and column are owned by different ranks makes the column a ghost of the row's owner."""
    n = part.n
    pos = np.empty(n, np.int64)
    pos[order] = np.arange(n)
    owner_of_pos = part.owner(np.arange(n)).astype(np.int16)
    lens = np.diff(A_nat.row_ptr)
    reader = np.repeat(owner_of_pos[pos], lens)                  # owner of the row of every entry
    cpos = pos[A_nat.col_idx]
    src = owner_of_pos[cpos]
    m = reader != src
    reader, src, cpos = reader[m], src[m], cpos[m]
    out = {r: {0: {}, 1: {}} for r in range(part.world)}
    if len(cpos):
        key = (reader.astype(np.int64) * part.world + src) * 2 + (cpos >= part.nF)
        o = np.lexsort((cpos, key))
        key, cpos = key[o], cpos[o]
        keep = np.ones(len(key), bool)
        keep[1:] = (key[1:] != key[:-1]) | (cpos[1:] != cpos[:-1])
        key, cpos = key[keep], cpos[keep]
        cuts = np.flatnonzero(np.concatenate(([True], key[1:] != key[:-1], [True])))
        for a, b in zip(cuts[:-1], cuts[1:]):
            k = int(key[a])
            which, rs = k & 1, k >> 1
            out[rs // part.world][which][rs % part.world] = cpos[a:b]
    return out


def interior_split(item_flags, a, b):
    """items [a, b) of a pass, item_flags[it] = the item has a row that reads a ghost entry.  Returns (interior, boundary): the longest run of
    unflagged items (they can run while the halo exchange is in flight) and the at most two ranges around it (after the wait)."""
    if b <= a:
        return (a, a), []
    f = np.asarray(item_flags[a:b], dtype=bool)
    best, run0 = (0, 0), None
    for i, v in enumerate(np.append(f, True)):
        if not v and run0 is None:
            run0 = i
        elif v and run0 is not None:
            if i - run0 > best[1] - best[0]:
                best = (run0, i)
            run0 = None
    i0, i1 = a + best[0], a + best[1]
    return (i0, i1), [r for r in ((a, i0), (i1, b)) if r[1] > r[0]]


def all_ghost_lists_native(A_nat, order, part, reads_ghost=None):
    """the same lists from libamgb200.so's host helper (amgb200_ghost_lists: one OpenMP pass; the numpy version above takes 1.3 s at 256^3);
    reads_ghost: optional uint8 array of n entries, set to 1 for every schedule row that reads an entry owned by another rank"""
    import ctypes as C
    from . import capi
    L = capi.lib()
    world = part.world
    fb = np.ascontiguousarray(part.f_bounds, dtype=np.int64)
    cb = np.ascontiguousarray(part.c_bounds, dtype=np.int64)
    order32 = np.ascontiguousarray(order, dtype=np.int32)
    ptr = np.zeros(world * world * 2 + 1, np.int64)
    idx = C.POINTER(C.c_int)()
    L.amgb200_ghost_lists_ex.restype = C.c_longlong
    L.amgb200_ghost_lists_ex.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.POINTER(C.POINTER(C.c_int)), C.c_void_p, C.c_void_p]
    total = L.amgb200_ghost_lists_ex(C.byref(A_nat.c), order32.ctypes.data, int(part.nF), int(world), fb.ctypes.data, cb.ctypes.data, C.byref(idx), ptr.ctypes.data,
                                     reads_ghost.ctypes.data if reads_ghost is not None else None)
    flat = np.ctypeslib.as_array(idx, shape=(max(int(total), 1),))[:int(total)].astype(np.int64)
    C.CDLL(None).free(idx)
    out = {r: {0: {}, 1: {}} for r in range(world)}
    for key in range(world * world * 2):
        a, b = int(ptr[key]), int(ptr[key + 1])
        if b > a:
            which, rs = key & 1, key >> 1
            out[rs // world][which][rs % world] = flat[a:b]
    return out


class ShardedSolver:
    """V-cycle solve with level 0 sharded over `world` ranks (see module docstring).

    backend interface (all vectors in level-0 schedule numbering):
      shape() -> dict(n, nF, itemsF, itemsC, rows_per_item, p_items, shardable)
      order() -> schedule->natural permutation of level 0
      gs_pass(which, item0, item1); residual(item0, item1); prolong(item0, item1)
      x0(), wp0(), x1()            -> writable 1-D torch tensors aliasing the level vectors
      restrict_and_lower_levels()  -> b1 = R wp0, x1 = 0, V-cycle on levels >= 1   (rank 0)
      set_problem(x_nat, b_nat), get_solution() ; sumsq(tensor) -> float
    """

    def __init__(self, backend, A_nat, dist, rank, world, pre=2, post=2):
        import torch
        self.torch, self.dist, self.be, self.rank, self.world = torch, dist, backend, rank, world
        sh = backend.shape()
        if not sh["shardable"]:
            raise ValueError("level 0 is not two-colour: the path does not shard (run replicas instead)")
        self.sh = sh
        self.part = Partition(sh["n"], sh["nF"], sh["itemsF"], sh["itemsC"], sh["rows_per_item"], world)
        self.pre, self.post = pre, post
        order = backend.order()
        # what I need from each peer, and (same function evaluated for the peer) what each peer needs from me
        native = hasattr(A_nat, "c") and getattr(backend, "native_lists", False)
        reads_ghost = np.zeros(sh["n"], np.uint8) if native else None
        allg = all_ghost_lists_native(A_nat, order, self.part, reads_ghost) if native else all_ghost_lists(A_nat, order, self.part)
        mine = allg[rank]
        self.recv_idx = {w: {src: torch.as_tensor(v, dtype=torch.long, device=backend.device) for src, v in mine[w].items()} for w in (0, 1)}
        self.send_idx = {0: {}, 1: {}}
        for peer in range(world):
            if peer == rank:
                continue
            theirs = allg[peer]
            for w in (0, 1):
                if rank in theirs[w]:
                    self.send_idx[w][peer] = torch.as_tensor(theirs[w][rank], dtype=torch.long, device=backend.device)
        self.halo_bytes = sum(8 * len(v) for w in (0, 1) for v in self.recv_idx[w].values())
        # prolongation item ranges (P's items are 32-row groups counted from schedule row 0) covering my rows;
        # at the seams an item may also touch rows of a neighbour: those are ghosts here, refreshed before use
        self.peer = hasattr(backend, "peer_setup") and world > 1 and not getattr(backend, "no_peer", False)
        if self.peer:
            # peer-memory plans (NVLink P2P stores + flag words, no collective, no host round trip): 0/1 ghost entries of the F/C pass,
            # 2 fine residual -> rank 0, 3 coarse correction rank 0 -> all, 4 solution -> rank 0
            own = [(a, b) for a, b in (self.part.f_rows[rank], self.part.c_rows[rank]) if b > a]
            plans = {}
            for w in (0, 1):
                plans[w] = {"vec": "x0", "sends": [(peer, idx.cpu().numpy().astype(np.int32), None) for peer, idx in self.send_idx[w].items()],
                            "srcs": sorted(self.recv_idx[w].keys())}
            for pid, vec in ((2, "wp0"), (4, "x0")):
                plans[pid] = {"vec": vec, "sends": [(0, None, r) for r in own] if rank else [], "srcs": list(range(1, world)) if rank == 0 else []}
            plans[3] = {"vec": "x1", "sends": [(peer, None, (0, sh.get("n1", 0))) for peer in range(1, world)] if rank == 0 else [],
                        "srcs": [0] if rank else []}
            backend.peer_setup(dist, rank, world, plans)
        rpi = sh["rows_per_item"]
        # interior / boundary split of my items of each pass (peer mode): the interior items read no ghost entry and run while the exchange
        # is in flight, the boundary items after the wait
        self.split = None
        import os
        if self.peer and reads_ghost is not None and hasattr(backend, "peer_start") and not int(os.environ.get("AMGB200_NO_OVERLAP", "0")):
            nF, n = sh["nF"], sh["n"]
            def item_flags(rows):                                 # rows of one pass -> flag per item of rpi rows
                pad = (-len(rows)) % rpi
                return np.pad(rows, (0, pad)).reshape(-1, rpi).any(axis=1)
            flags = {0: item_flags(reads_ghost[:nF]), 1: item_flags(reads_ghost[nF:n])}
            self.split = {0: interior_split(flags[0], *self.part.f_items[rank]), 1: interior_split(flags[1], *self.part.c_items[rank])}
        self.p_ranges = []
        for a, b in (self.part.f_rows[rank], self.part.c_rows[rank]):
            if b > a:
                lo, hi = a // rpi, min(sh["p_items"], -(-b // rpi))
                if self.p_ranges and lo < self.p_ranges[-1][1]:   # never run an item twice on one rank
                    lo = self.p_ranges[-1][1]
                if hi > lo:
                    self.p_ranges.append((lo, hi))

    # ---- communication -------------------------------------------------------------------------
    def exchange(self, which):
        """refresh my ghost copies of pass-`which` x entries from their owners (point-to-point)"""
        if self.world == 1:
            return
        if self.peer:
            self.be.peer_run(which)
            return
        torch, dist = self.torch, self.dist
        x = self.be.x0()
        ops, recv_bufs = [], []
        for peer, idx in self.send_idx[which].items():
            ops.append(dist.P2POp(dist.isend, x.index_select(0, idx).contiguous(), peer))
        for peer, idx in self.recv_idx[which].items():
            buf = torch.empty(len(idx), dtype=x.dtype, device=x.device)
            recv_bufs.append((idx, buf))
            ops.append(dist.P2POp(dist.irecv, buf, peer))
        if ops:
            for req in dist.batch_isend_irecv(ops):
                req.wait()
        for idx, buf in recv_bufs:
            x.index_copy_(0, idx, buf)

    def gather_residual(self):
        """own rows of wp0 -> rank 0 (the restriction needs the whole fine residual)"""
        if self.world == 1:
            return
        if self.peer:
            self.be.peer_run(2)
            return
        torch, dist = self.torch, self.dist
        wp = self.be.wp0()
        ops = []
        if self.rank == 0:
            for peer in range(1, self.world):
                for a, b in (self.part.f_rows[peer], self.part.c_rows[peer]):
                    if b > a:
                        ops.append(dist.P2POp(dist.irecv, wp[a:b], peer))
        else:
            for a, b in (self.part.f_rows[self.rank], self.part.c_rows[self.rank]):
                if b > a:
                    ops.append(dist.P2POp(dist.isend, wp[a:b].contiguous(), 0))
        if ops:
            for req in dist.batch_isend_irecv(ops):
                req.wait()

    # ---- one V-cycle ----------------------------------------------------------------------------
    def smooth(self, sweeps):
        fa, fb = self.part.f_items[self.rank]
        ca, cb = self.part.c_items[self.rank]
        if self.split is not None:
            # halo exchange overlapped with the interior rows: start (stores + flags), interior items, wait, boundary items
            for _ in range(sweeps):
                for which in (0, 1):                     # the F pass reads C ghosts (plan 1), the C pass reads F ghosts (plan 0)
                    (i0, i1), boundary = self.split[which]
                    self.be.peer_start(1 - which)
                    if i1 > i0:
                        self.be.gs_pass(which, i0, i1)
                    self.be.peer_wait(1 - which)
                    for a, b in boundary:
                        self.be.gs_pass(which, a, b)
            return
        for _ in range(sweeps):
            self.exchange(1)                 # F rows read C neighbours
            self.be.gs_pass(0, fa, fb)
            self.exchange(0)                 # C rows read F neighbours
            self.be.gs_pass(1, ca, cb)

    def residual_own(self):
        self.exchange(1)                     # F values are current since the last C pass; C values changed
        self.exchange(0)
        fa, fb = self.part.f_items[self.rank]
        ca, cb = self.part.c_items[self.rank]
        self.be.residual(fa, fb)
        self.be.residual(self.sh["itemsF"] + ca, self.sh["itemsF"] + cb)

    def cycle(self):
        self.smooth(self.pre)
        self.residual_own()
        self.gather_residual()
        if self.rank == 0:
            self.be.restrict_and_lower_levels()
        if self.peer:
            self.be.peer_run(3)                           # coarse correction stored into every rank's x1
        elif self.world > 1:
            self.dist.broadcast(self.be.x1(), src=0)      # coarse correction to every rank
        for a, b in self.p_ranges:                        # own rows (+ partial items at the seams)
            self.be.prolong(a, b)
        self.smooth(self.post)

    def residual_norm(self):
        self.residual_own()
        wp = self.be.wp0()
        s = self.torch.zeros(1, dtype=wp.dtype, device=wp.device)
        for a, b in (self.part.f_rows[self.rank], self.part.c_rows[self.rank]):
            if b > a:
                s += (wp[a:b] * wp[a:b]).sum()
        if self.world > 1:
            self.dist.all_reduce(s)
        return float(s.sqrt().item())

    def collect_solution(self):
        """owners -> rank 0 (level-0 x in schedule numbering)"""
        if self.world == 1:
            return
        if self.peer:
            self.be.peer_run(4)
            return
        dist = self.dist
        x = self.be.x0()
        ops = []
        if self.rank == 0:
            for peer in range(1, self.world):
                for a, b in (self.part.f_rows[peer], self.part.c_rows[peer]):
                    if b > a:
                        ops.append(dist.P2POp(dist.irecv, x[a:b], peer))
        else:
            for a, b in (self.part.f_rows[self.rank], self.part.c_rows[self.rank]):
                if b > a:
                    ops.append(dist.P2POp(dist.isend, x[a:b].contiguous(), 0))
        if ops:
            for req in dist.batch_isend_irecv(ops):
                req.wait()

    def solve(self, x_nat, b_nat, tol, max_it=100):
        """outer iteration of SSS_amg_solve (Solve/SSS_SOLVE.c:53-80); returns (nits, history, x on rank 0)"""
        self.be.set_problem(x_nat, b_nat)
        sumb = float(np.sqrt(np.sum(np.asarray(b_nat, dtype=np.float64) ** 2)))
        hist = []
        for it in range(1, max_it + 1):
            self.cycle()
            absres = self.residual_norm()
            hist.append(absres)
            if absres / sumb < tol:
                break
        self.collect_solution()
        return len(hist), np.array(hist), (self.be.get_solution() if self.rank == 0 else None)

    def solve_resident(self, d_x0, d_b, sumb, tol, d_out, max_it=100):
        """the same outer iteration on vectors that already live on every rank's GPU (torch tensors, natural numbering); the solution
        stays on the device (rank 0: d_out).  Nothing crosses PCIe except the 8-byte norm of every V-cycle."""
        self.be.set_problem(d_x0, d_b)
        hist = []
        for it in range(1, max_it + 1):
            self.cycle()
            absres = self.residual_norm()
            hist.append(absres)
            if absres / sumb < tol:
                break
        self.collect_solution()
        if self.rank == 0:
            self.be.get_solution_device(d_out)
        return len(hist), np.array(hist)


class _CudaArray:
    """zero-copy torch view of a device pointer owned by libamgb200 (via __cuda_array_interface__)"""

    def __init__(self, ptr, n):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": "<f8", "data": (int(ptr), False), "version": 2}


class GpuBackend:
    """libamgb200.so building blocks (include/amg_b200.h section 2b) on this rank's GPU"""

    def __init__(self, dev, torch):
        import ctypes as C
        from . import capi
        self.C, self.capi, self.torch, self.dev = C, capi, torch, dev
        self.L = capi.lib()
        self.h = dev.h
        self.device = torch.device("cuda", torch.cuda.current_device())
        self.L.amgb200_set_stream(self.h, C.c_void_p(torch.cuda.current_stream().cuda_stream))
        info = (C.c_longlong * 8)()
        self.L.amgb200_l0_shape(self.h, info)
        self._shape = {"n": info[0], "nF": info[1], "itemsF": info[2], "itemsC": info[3], "rows_per_item": info[6],
                       "p_items": info[7], "shardable": bool(info[5]) and dev.num_levels >= 2}
        n0, n1 = info[0], dev.info(1)["rows"] if dev.num_levels > 1 else 0
        self._shape["n1"] = n1
        import os
        self.native_lists = True                                               # ghost lists from the library's host helper
        self.no_peer = bool(int(os.environ.get("AMGB200_NO_PEER", "0")))      # (developer switch: NCCL point-to-point instead of peer stores)
        self._x0 = torch.as_tensor(_CudaArray(self.L.amgb200_level_vec(self.h, 0, 0), n0), device=self.device)
        self._b0 = torch.as_tensor(_CudaArray(self.L.amgb200_level_vec(self.h, 0, 1), n0), device=self.device)
        self._wp0 = torch.as_tensor(_CudaArray(self.L.amgb200_level_vec(self.h, 0, 2), n0), device=self.device)
        self._x1 = torch.as_tensor(_CudaArray(self.L.amgb200_level_vec(self.h, 1, 0), n1), device=self.device) if n1 else None
        self._nat = torch.empty(n0, dtype=torch.float64, device=self.device)

    def shape(self):
        return self._shape

    # ---- peer-memory exchange plans (include/amg_b200.h: amgb200_ipc_*, amgb200_peer_*) ----
    def peer_setup(self, dist, rank, world, plans):
        C, L = self.C, self.L
        vecs = {"x0": (0, 0), "wp0": (0, 2), "x1": (1, 0), "flags": (0, 3)}
        mine = {}
        for name, (lvl, which) in vecs.items():
            buf = C.create_string_buffer(64)
            L.amgb200_ipc_export(self.h, lvl, which, buf)
            mine[name] = buf.raw
        everyone = [None] * world
        dist.all_gather_object(everyone, mine)
        opened = {}
        def peer_ptr(peer, name):
            if (peer, name) not in opened:
                opened[(peer, name)] = L.amgb200_ipc_open(self.h, everyone[peer][name])
            return opened[(peer, name)]
        local = {"x0": self._x0.data_ptr(), "wp0": self._wp0.data_ptr(), "x1": self._x1.data_ptr() if self._x1 is not None else 0}
        self._keep = []
        for pid, pl in sorted(plans.items()):
            sends = pl["sends"]
            n = len(sends)
            peer_vec = (C.c_void_p * max(n, 1))()
            idx = (self.capi.c_int_p * max(n, 1))()
            count = (C.c_int * max(n, 1))()
            range0 = (C.c_int * max(n, 1))()
            flag_peers = sorted({peer for peer, _, _ in sends})
            for i, (peer, ix, rng) in enumerate(sends):
                peer_vec[i] = peer_ptr(peer, pl["vec"])
                if ix is not None:
                    arr = np.ascontiguousarray(ix, np.int32)
                    self._keep.append(arr)
                    idx[i] = self.capi.iptr(arr); count[i] = len(arr); range0[i] = 0
                else:
                    idx[i] = self.capi.c_int_p(); count[i] = rng[1] - rng[0]; range0[i] = rng[0]
            flags = (C.c_void_p * max(len(flag_peers), 1))()
            for i, peer in enumerate(flag_peers):
                flags[i] = peer_ptr(peer, "flags") + 4 * (pid * 64 + rank)
            srcs = (C.c_int * max(len(pl["srcs"]), 1))(*pl["srcs"]) if pl["srcs"] else (C.c_int * 1)()
            L.amgb200_peer_plan(self.h, pid, n, C.c_void_p(local[pl["vec"]]), peer_vec, idx, count, range0, len(flag_peers), flags,
                                len(pl["srcs"]), srcs)
        dist.barrier()        # every rank's plans (and flag words) exist before anybody runs one

    def peer_run(self, plan):
        self.L.amgb200_peer_run(self.h, plan)

    def peer_start(self, plan):
        self.L.amgb200_peer_start(self.h, plan)

    def peer_wait(self, plan):
        self.L.amgb200_peer_wait(self.h, plan)

    def order(self):
        o = np.zeros(self._shape["n"], np.int32)
        self.L.amgb200_level_order(self.h, 0, self.capi.iptr(o))
        return o.astype(np.int64)

    def x0(self): return self._x0
    def wp0(self): return self._wp0
    def x1(self): return self._x1
    def gs_pass(self, which, a, b): self.L.amgb200_l0_gs_pass(self.h, which, a, b)
    def residual(self, a, b): self.L.amgb200_l0_residual(self.h, a, b)
    def prolong(self, a, b): self.L.amgb200_l0_prolong(self.h, a, b)

    def restrict_and_lower_levels(self):
        self.L.amgb200_restrict_from(self.h, 0)
        self.L.amgb200_cycle_from(self.h, 1)

    def set_problem(self, x_nat, b_nat):
        """x_nat, b_nat: host arrays, or torch tensors already resident on this rank's GPU (natural numbering)"""
        t = self.torch
        for src, dst in ((x_nat, self._x0), (b_nat, self._b0)):
            if t.is_tensor(src):
                self._nat.copy_(src)
            else:
                self._nat.copy_(t.as_tensor(np.asarray(src, dtype=np.float64)))
            self.L.amgb200_vec_to_schedule(self.h, 0, self.C.c_void_p(self._nat.data_ptr()), self.C.c_void_p(dst.data_ptr()))

    def get_solution_device(self, out):
        """level-0 x in natural numbering into the resident tensor `out` (no host copy)"""
        self.L.amgb200_vec_to_natural(self.h, 0, self.C.c_void_p(self._x0.data_ptr()), self.C.c_void_p(out.data_ptr()))
        return out

    def get_solution(self):
        self.L.amgb200_vec_to_natural(self.h, 0, self.C.c_void_p(self._x0.data_ptr()), self.C.c_void_p(self._nat.data_ptr()))
        return self._nat.cpu().numpy()
