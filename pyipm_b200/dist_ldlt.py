"""2-D block-cyclic LDL^T of a dense symmetric (quasi-definite / KKT) matrix across the GPUs of one node.

BASELINE.json config 4 / SURVEY.md section 8(e): the dense KKT solve shards only when the matrix order makes a
panel factorisation natural (Kc >~ 16k).  One process per GPU, `torch.distributed` (NCCL over NVLink/NVSwitch)
for the plumbing, the compute is this repo's own kernels through the C ABI (tile-pivoted Bunch-Kaufman tile
factorisation, fused DMMA panel kernel, DMMA trailing update) -- the same building blocks as the single-GPU
factorisation (`pyipm_b200/csrc/ldlt.cuh`), so pivoting stays inside 64 x 64 diagonal tiles and no pivot search
ever crosses a GPU boundary.

Layout: blocks of `b = 256` rows/cols; block (I, J) lives on rank (I mod P, J mod Q) of a P x Q process grid
(rank = p * Q + q).  Every rank keeps the original local blocks (for the refinement residual) and a working copy.

Per block column k:
  1. the owner of (k, k) factors the diagonal block (4 tile steps) and broadcasts its factor data (L_kk, the
     per-tile L^-1 P, D^-1 blocks: 0.66 MB);
  2. the ranks of process column k mod Q turn their row blocks of the panel into L (in place) and W = L * D;
  3. the panel (L and W, <= 67 MB at n = 16384) is broadcast from its P owners to every rank -- NVSwitch gives each
     GPU full bandwidth to every peer, so the panel is simply replicated instead of routed along grid rows/cols;
  4. every rank updates its own trailing blocks  A[I, J] -= W[I] * L[J]^T  (I >= J) with K = 256 DMMA launches.
Inertia = sum of the diagonal-block counts.  The O(n^2) triangular solves are latency bound, so they are not
distributed: the factor is already replicated by step 3, each rank adopts it (`b200ipm_ldlt_import`) and runs the
single-launch flag-synchronised solves locally; iterative refinement uses the distributed original matrix
(local mat-vec + one all-reduce).

The tile arithmetic is behind a small `ops` object so that the block-cyclic indexing / collective logic can be
exercised on CPU (gloo, world_size 2) with a reference implementation (tests/test_dist_ldlt_cpu.py).
"""
from __future__ import print_function

import ctypes as C
import os

import numpy as np
import torch
import torch.distributed as dist

NB = 64


class CudaTileOps(object):
    """Tile arithmetic on the local GPU through libb200ipm.so (device pointers of torch tensors)."""

    def __init__(self, device, block=256):
        from . import _lib
        assert block % NB == 0
        self._lib = _lib
        self.lib = _lib.load()
        self.b = block
        self.nt = block // NB
        self.device = torch.device('cuda', device)
        stream = _lib.torch_stream_handle(self.device)
        self.ctx = _lib.DenseLDLT(block, device=device, stream=stream)  # carries device + stream for the tile calls
        # the serial chain (diagonal factor, panel, broadcasts) runs on its own high-priority stream so that it overlaps
        # with the trailing update of the previous block column (look-ahead 1)
        self.s_chain = torch.cuda.Stream(device=self.device, priority=-1)
        self.ctx_chain = _lib.DenseLDLT(block, device=device, stream=self.s_chain.cuda_stream)
        self.tile_doubles = NB * NB + 4 * NB + NB // 2                   # LinvP + [dinv_a|dinv_b|d_a|d_b] + kind (ints)
        self.diag_size = self.b * self.b + self.nt * self.tile_doubles + 3
        self.perm = torch.empty(NB, dtype=torch.int32, device=self.device)
        self.wdiag = torch.empty((self.b, self.b), dtype=torch.float64, device=self.device)
        self.wdiag_chain = torch.empty((self.b, self.b), dtype=torch.float64, device=self.device)
        self._idx_cache = {}

    def empty(self, *shape):
        return torch.empty(shape, dtype=torch.float64, device=self.device)

    def zeros(self, *shape):
        return torch.zeros(shape, dtype=torch.float64, device=self.device)

    def _tile(self, diag, t):
        off = self.b * self.b + t * self.tile_doubles
        return diag[off:off + NB * NB], diag[off + NB * NB:off + self.tile_doubles]

    # ---- stream plumbing of the look-ahead pipeline
    def chain(self):
        return torch.cuda.stream(self.s_chain)

    def record(self):
        e = torch.cuda.Event()
        e.record(torch.cuda.current_stream(self.device))
        return e

    def wait(self, ev):
        if ev is not None:
            torch.cuda.current_stream(self.device).wait_event(ev)

    def record_timed(self):
        e = torch.cuda.Event(enable_timing=True)
        e.record(torch.cuda.current_stream(self.device))
        return e

    def sync(self):
        torch.cuda.synchronize(self.device)

    def factor_diag(self, Akk, diag, chain=False):
        """Akk: b x b view (row stride ld) of the working matrix; factored in place; fills `diag` (one C call)."""
        ctx = self.ctx_chain if chain else self.ctx
        wd = self.wdiag_chain if chain else self.wdiag
        self._lib.check(self.lib.b200ipm_ldlt_block_factor(ctx.h, Akk.data_ptr(), Akk.stride(0), self.b, diag.data_ptr(),
                                                           wd.data_ptr()))
        # (inertia is evaluated once, for all diagonal blocks together, by counts_all)

    def panel(self, Bblk, diag, chain=False, out=None):
        """Bblk: rows x b view (row stride ld) -> overwritten with L; returns W = L * D (rows x b, contiguous; written
        into `out` when given)."""
        ctx = self.ctx_chain if chain else self.ctx
        W = self.empty(Bblk.shape[0], self.b) if out is None else out
        self._lib.check(self.lib.b200ipm_ldlt_block_panel(ctx.h, Bblk.data_ptr(), Bblk.stride(0), Bblk.shape[0], self.b,
                                                          diag.data_ptr(), W.data_ptr()))
        return W

    def colblock(self, Acol, diag, Wb, chain=True, rest_event=None):
        """Acol: (rows_total x b) view of a whole block column from its diagonal block down (row stride ld): factored in
        place with the single-GPU panel schedule (one C call); Wb (rows_total x b, contiguous) <- W = L D; fills `diag`.
        rest_event: the rows below the diagonal block are not touched before this (recorded) event has fired."""
        ctx = self.ctx_chain if chain else self.ctx
        ev = C.c_void_p(rest_event.cuda_event) if rest_event is not None else None
        self._lib.check(self.lib.b200ipm_ldlt_colblock_factor(ctx.h, Acol.data_ptr(), Acol.stride(0), Acol.shape[0], self.b,
                                                              diag.data_ptr(), Wb.data_ptr(), ev))

    # ---- tcgen05 trailing updates (int8 error-free split, the kernels of the single-GPU factorisation)
    def tc_slice(self, W, L, max_rows):
        """digits of W (rows x b) and -L (rows x b) of the current panel, once per panel"""
        self._lib.check(self.lib.b200ipm_oz_panel_slice(self.ctx.h, W.shape[0], W.data_ptr(), W.stride(0), L.data_ptr(),
                                                        L.stride(0), int(max_rows)))

    def tc_update(self, Cv, row_off):
        """Cv (n x ncols lower trapezoid with its origin on the diagonal; its first row is row `row_off` of the sliced
        panel) -= W L^T, in place"""
        self._lib.check(self.lib.b200ipm_oz_block_update(self.ctx.h, Cv.data_ptr(), Cv.stride(0), Cv.shape[0], Cv.shape[1],
                                                         int(row_off)))

    def tc_status(self):
        err = C.c_int(0)
        self._lib.check(self.lib.b200ipm_oz_status(self.ctx.h, C.byref(err)))
        return err.value

    def update(self, Cv, W, L):
        """Cv (rows x cols view, row stride ldc) -= W (rows x b) @ L (cols x b)^T"""
        rows, cols = Cv.shape
        if rows == 0 or cols == 0:
            return
        self._lib.check(self.lib.b200ipm_gemm_nt_update(self.ctx.h, Cv.data_ptr(), Cv.stride(0), rows, cols, W.data_ptr(),
                                                        W.stride(0), L.data_ptr(), L.stride(0), self.b, 0))

    def update_bc(self, Cv, W, L, grid, coord, li0, lj0):
        """One launch: Cv (local rows x local cols from local block (li0, lj0)) -= W @ L^T on the tiles whose global
        block row >= global block column."""
        rows, cols = Cv.shape
        if rows == 0 or cols == 0:
            return
        self._lib.check(self.lib.b200ipm_gemm_nt_update_bc(self.ctx.h, Cv.data_ptr(), Cv.stride(0), rows, cols, W.data_ptr(),
                                                           W.stride(0), L.data_ptr(), L.stride(0), self.b, self.b, grid[0],
                                                           grid[1], coord[0], coord[1], li0, lj0))

    def matvec_local(self, A0, Xc):
        """Xc (nrhs x cols, contiguous) -> nrhs x rows:  row r = A0 * Xc[r]  with the repo's GEMV kernel"""
        rows, cols = A0.shape
        out = self.empty(Xc.shape[0], rows)
        for r in range(Xc.shape[0]):
            self._lib.check(self.lib.b200ipm_ldlt_gemv(self.ctx.h, A0.data_ptr(), A0.stride(0), rows, cols,
                                                       Xc[r].data_ptr(), out[r].data_ptr()))
        return out

    def counts_all(self, diags):
        """(pos, neg, zero) over all diagonal blocks: signs of the 1x1 / 2x2 blocks of D, one batched evaluation and
        one synchronisation."""
        NT = self.tile_doubles
        tiles = torch.stack([d[self.b * self.b:self.b * self.b + self.nt * NT].view(self.nt, NT) for d in diags])
        tiles = tiles.reshape(-1, NT)[:, NB * NB:]                      # [ntiles, 288]: dinv_a | dinv_b | d_a | d_b | kind
        da, db = tiles[:, 2 * NB:3 * NB], tiles[:, 3 * NB:4 * NB]
        kind = tiles[:, 4 * NB:].contiguous().view(torch.int32)[:, :NB]
        one, first = (kind == 0), (kind == 1)
        c = torch.roll(da, -1, dims=1)                                   # second diagonal entry of a 2x2 block
        det, tr = da * c - db * db, da + c
        pos = (one & (da > 0)).sum() + (first & (det > 0) & (tr > 0)).sum() * 2 + (first & (det < 0)).sum()
        neg = (one & (da < 0)).sum() + (first & (det > 0) & (tr < 0)).sum() * 2 + (first & (det < 0)).sum()
        zero = (one & (da == 0)).sum() + (first & (det == 0)).sum()
        return [int(v) for v in torch.stack([pos, neg, zero]).tolist()]

    def index(self, idx):
        """device index tensor (int64) from a NumPy index array, cached: no per-panel host-to-device copies"""
        key = idx.tobytes()
        t = self._idx_cache.get(key)
        if t is None:
            t = torch.from_numpy(idx).to(self.device)
            self._idx_cache[key] = t
        return t

    # ---- replicated solve on the gathered factor
    def make_solver(self, n, diags, panels):
        """Replicated solver on the gathered factor.  The native handle and the staging buffers are created ONCE per order
        and re-used by later factorisations (creating / destroying a 10 GB workspace costs tens of milliseconds of
        synchronous cudaMalloc / cudaFree)."""
        b, nt = self.b, self.nt
        nblk = n // NB
        cache = getattr(self, '_solver_cache', None)
        if cache is None or cache[0] != n:
            stream = self._lib.torch_stream_handle(self.device)
            cache = (n, self.zeros(n, n), self.empty(nblk, NB, NB), self.zeros(4, n),
                     torch.zeros(n, dtype=torch.int32, device=self.device),
                     self._lib.DenseLDLT(n, device=self.device.index, stream=stream))
            self._solver_cache = cache
        _, A, linvp, dinfo, kind, F = cache
        for k, diag in enumerate(diags):
            r0 = k * b
            A[r0:r0 + b, r0:r0 + b] = diag[:b * b].view(b, b)
            if panels[k] is not None:
                A[r0 + b:, r0:r0 + b] = panels[k]
        # per-tile factor data of all diagonal blocks in three batched copies
        NT = self.tile_doubles
        tiles = torch.stack([d[b * b:b * b + nt * NT].view(nt, NT) for d in diags]).reshape(nblk, NT)
        linvp.copy_(tiles[:, :NB * NB].reshape(nblk, NB, NB))
        dinfo.copy_(tiles[:, NB * NB:NB * NB + 4 * NB].reshape(nblk, 4, NB).permute(1, 0, 2).reshape(4, n))
        kind.copy_(tiles[:, NB * NB + 4 * NB:].contiguous().view(torch.int32)[:, :NB].reshape(n))
        self._lib.check(self.lib.b200ipm_ldlt_import(F.h, A.data_ptr(), n, linvp.data_ptr(), dinfo.data_ptr(), kind.data_ptr()))
        lib, chk = self.lib, self._lib.check

        def solve(Bt):   # Bt: nrhs x n contiguous device tensor, solved in place
            chk(lib.b200ipm_ldlt_solve(F.h, Bt.data_ptr(), Bt.shape[0], 0, 1))
            return Bt
        solve.keepalive = F
        return solve


class BlockCyclicLDLT(object):
    def __init__(self, n, grid, ops, block=256, group=None):
        self.n, self.b, self.ops, self.group = int(n), int(block), ops, group
        self.P, self.Q = grid
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        assert self.P * self.Q == self.world, 'grid does not match the world size'
        assert self.n % self.b == 0
        self.nbk = self.n // self.b
        self.p, self.q = divmod(self.rank, self.Q)
        self.rows_blk = [I for I in range(self.nbk) if I % self.P == self.p]     # my global block rows
        self.cols_blk = [J for J in range(self.nbk) if J % self.Q == self.q]     # my global block cols
        self.A0 = None
        self.diags, self.panels, self.inertia = [], [], None

    # ---- helpers
    def _rank_of(self, p, q):
        return p * self.Q + q

    def _idx(self, blocks):
        b = self.b
        if not blocks:
            return np.zeros(0, dtype=np.int64)
        return np.concatenate([np.arange(I * b, (I + 1) * b) for I in blocks])

    def _bcast(self, t, src):
        if self.world > 1:
            dist.broadcast(t, src=src, group=self.group)

    def load(self, A):
        """A: full n x n symmetric matrix (NumPy, same on every rank).  Keeps the local blocks of BOTH triangles
        (the refinement mat-vec uses them); the factorisation touches blocks I >= J of a working copy."""
        ri, ci = self._idx(self.rows_blk), self._idx(self.cols_blk)
        loc = np.ascontiguousarray(A[np.ix_(ri, ci)])
        self.A0 = self.ops.empty(*loc.shape)
        self.A0.copy_(torch.from_numpy(loc))
        self.ri = torch.from_numpy(ri).to(self.A0.device)
        self.ci = torch.from_numpy(ci).to(self.A0.device)

    def load_device(self, A):
        """A: full n x n symmetric matrix as a device tensor (identical on every rank, e.g. generated from the
        same seed).  Only the local blocks are kept."""
        ri = torch.from_numpy(self._idx(self.rows_blk)).to(A.device)
        ci = torch.from_numpy(self._idx(self.cols_blk)).to(A.device)
        self.A0 = A.index_select(0, ri).index_select(1, ci).contiguous()
        self.ri, self.ci = ri, ci

    def factor(self):
        """-> inertia (pos, neg, zero).  Collective: every rank must call it.  1 x Q grids take the block-column-cyclic
        pipeline (`_factor_cols`), P > 1 the general 2-D one below."""
        if self.P == 1:
            return self._factor_cols()
        return self._factor_2d()

    def _factor_cols(self):
        """1 x Q grid = 1-D block-column-cyclic: block column k lives whole on rank k mod Q, so the serial part of a step --
        diagonal block (4 tile steps) and the panel below it -- never leaves one GPU, and a step has exactly ONE collective:
        the broadcast of the contiguous record [factor data of the diagonal block | L | W = L D] from the owner (NVSwitch:
        every peer at full bandwidth).  Look-ahead 1 on two streams: the UPDATE stream of the owner of column k + 1 updates
        that block column first; its CHAIN stream (high priority) then factors it and broadcasts while every rank is still
        applying panel k to the rest of its block columns.  The records are persistent (one buffer per block column, reused
        by later factorisations): no ring, no reshuffling copies; they are also what the replicated solver imports."""
        ops, b, Q, nbk = self.ops, self.b, self.Q, self.nbk
        ds = ops.diag_size
        dsp = (ds + 31) // 32 * 32     # L and W start on 256-byte boundaries (vectorised / TMA-staged kernels read them)
        fused = hasattr(ops, 'colblock')     # CUDA backend: diagonal block + panel in one call (single-GPU panel schedule)
        # ... and the trailing updates beyond the look-ahead column on tcgen05 (B200IPM_DIST_TC=0: fp64 DMMA)
        use_tc = hasattr(ops, 'tc_slice') and os.environ.get('B200IPM_DIST_TC', '1') != '0'
        tc_used = False
        # With >= 8 ranks the update stream has slack (a rank's share of a trailing update is shorter than a step of the
        # chain), so the owner of the NEXT block column postpones its own bulk update until its chain work is on its way:
        # the column-block factorisation then does not share the SMs with a saturating update (0.36 -> 0.22 ms per column).
        d_env = os.environ.get('B200IPM_DIST_DEFER')
        defer = (Q >= 8) if d_env is None else (d_env != '0')
        deferred = None
        if getattr(self, '_pan', None) is None:
            # record of block column k: [factor data of the diagonal block | L (rows x b) | b x b scratch (W of the diagonal
            # block's own rows) | W (rows x b)], contiguous: one broadcast
            self._pan = [ops.empty(dsp + (2 * (nbk - 1 - k) + 1) * b * b) for k in range(nbk)]
            self._work = ops.empty(*self.A0.shape)
        work = self._work
        work.copy_(self.A0)
        self.diags, self.panels, self._pgrp = [], [], None
        ev_look = ops.record()         # block column 0 is ready: `work` is a fresh copy made on the update stream
        ev_rest = None                 # ... and so are its rows below the diagonal block
        prof = [] if getattr(self, 'profile', False) else None     # per-column timing events (tools/prof_dist.py)
        tev = (lambda: ops.record_timed()) if prof is not None else (lambda: None)
        for k in range(nbk):
            owner = k % Q
            rows = (nbk - 1 - k) * b
            buf = self._pan[k]
            diag = buf[:ds]
            Lk = buf[dsp:dsp + rows * b].view(rows, b) if rows else None
            Wd = buf[dsp + rows * b:]                                   # (b + rows) x b: W of the whole block column
            Wk = Wd[b * b:].view(rows, b) if rows else None
            with ops.chain():
                t0 = t1 = t2 = None
                if owner == self.q:
                    ops.wait(ev_look)
                    t0 = tev()
                    lj = k // Q
                    if not fused:
                        ops.wait(ev_rest)
                    if fused:
                        ops.colblock(work[k * b:, lj * b:(lj + 1) * b], diag, Wd, chain=True, rest_event=ev_rest)
                        t1 = tev()
                        if rows:
                            Lk.copy_(work[(k + 1) * b:, lj * b:(lj + 1) * b])
                    else:
                        ops.factor_diag(work[k * b:(k + 1) * b, lj * b:(lj + 1) * b], diag, chain=True)
                        t1 = tev()
                        if rows:
                            Bv = work[(k + 1) * b:, lj * b:(lj + 1) * b]
                            ops.panel(Bv, diag, chain=True, out=Wk)
                            Lk.copy_(Bv)
                    t2 = tev()
                tb0 = tev()
                self._bcast(buf, self._rank_of(0, owner))
                ev_panel = ops.record()
                tb1 = tev()
            self.diags.append(diag)
            self.panels.append(Lk)
            if not rows:
                break
            # trailing update of my block columns J > k:  A[J.., J] -= W[J..] L[J]^T, block column k + 1 first
            ops.wait(ev_panel)
            if deferred is not None:       # my bulk update of the previous panel, held back while I factored this column
                deferred()
                deferred = None
            ev_look = None
            mycols = [J for J in self.cols_blk if J > k]
            tu0 = tev()
            tu1 = None
            if mycols:
                lj0 = mycols[0] // Q
                rest = mycols
                if mycols[0] == k + 1:
                    # the next block column first, its diagonal block before the rows below it: the owner's chain stream
                    # starts the serial tile steps as soon as the diagonal block is up to date
                    col = work[(k + 1) * b:, lj0 * b:(lj0 + 1) * b]
                    ops.update_bc(col[:b], Wk[:b], Lk[:b], (1, Q), (0, self.q), k + 1, lj0)
                    ev_look = ops.record()
                    ev_rest = None
                    if rows > b:
                        ops.update_bc(col[b:], Wk[b:], Lk[:b], (1, Q), (0, self.q), k + 2, lj0)
                        ev_rest = ops.record()
                    tu1 = tev()
                    rest, lj0 = mycols[1:], lj0 + 1
                def bulk(k=k, rest=rest, lj0=lj0, Wk=Wk, Lk=Lk):
                    if use_tc:
                        # digits of the panel once, then one in-place tcgen05 update per owned block column (a single
                        # launch over the whole trailing trapezoid when this rank owns every column)
                        ops.tc_slice(Wk, Lk, (nbk - 1) * b)
                        if Q == 1:
                            off = (rest[0] - k - 1) * b
                            ops.tc_update(work[rest[0] * b:, rest[0] * b:], off)
                        else:
                            for J in rest:
                                lj = J // Q
                                ops.tc_update(work[J * b:, lj * b:(lj + 1) * b], (J - k - 1) * b)
                    else:
                        if Q == 1:
                            Lm = Lk[(rest[0] - k - 1) * b:]
                        else:
                            pos = np.concatenate([np.arange(b) + (J - k - 1) * b for J in rest])
                            Lm = Lk.index_select(0, ops.index(pos))             # L rows of my block columns
                        ops.update_bc(work[(k + 1) * b:, lj0 * b:], Wk, Lm, (1, Q), (0, self.q), k + 1, lj0)
                if rest:
                    tc_used = tc_used or use_tc
                    if defer and mycols[0] == k + 1 and k + 2 < nbk:
                        deferred = bulk
                    else:
                        bulk()
            if ev_look is None:
                ev_look = ops.record()
            if prof is not None:
                prof.append((k, owner, t0, t1, t2, tb0, tb1, tu0, tu1, tev()))
        if deferred is not None:
            deferred()
        ops.sync()
        if tc_used and ops.tc_status():
            raise FloatingPointError('block-column-cyclic LDL^T: non-finite entries met by the tcgen05 trailing update')
        if prof is not None:
            self.profile_events = prof
        self.inertia = tuple(int(v) for v in ops.counts_all(self.diags))   # one synchronisation at the very end
        self._solver = None
        return self.inertia

    def _factor_2d(self):
        """General P x Q grid.

        Pipeline with look-ahead 1 on two streams.  CHAIN stream (high priority), per block column k: wait until column k
        carries every update through step k - 1; the owner of (k, k) factors the diagonal block and broadcasts its factor
        data; process column k mod Q computes its rows of the panel; the panel is broadcast from its P owners straight
        into per-owner slices of one buffer (no reshuffling copies).  UPDATE stream: as soon as panel k has arrived, the
        blocks of block column k + 1 are updated FIRST (that is all the chain's next step waits for), then everything to
        the right of it -- overlapped with the chain work and the broadcasts of step k + 1."""
        ops, b, P, Q = self.ops, self.b, self.P, self.Q
        work = self.A0.clone()
        self.diags, self.panels, self._pgrp = [], [], []
        if getattr(self, '_Wbuf', None) is None:
            self._Wbuf = [ops.empty(max(self.nbk - 1, 1) * b, b) for _ in range(2)]
        ev_look = None                 # block column k is up to date (recorded on the update stream)
        ev_wfree = [None, None]        # the W buffer of parity k % 2 is no longer read by an update
        ev_start = ops.record()        # `work` is a fresh clone made on the update stream
        for k in range(self.nbk):
            pk, qk = k % P, k % Q
            nbelow = self.nbk - (k + 1)
            groups = [[I for I in range(k + 1, self.nbk) if I % P == ps] for ps in range(P)]
            offs = np.concatenate([[0], np.cumsum([len(gp) * b for gp in groups])]).astype(np.int64)
            mine = groups[self.p]
            with ops.chain():
                ops.wait(ev_start if k == 0 else ev_look)
                # 1. diagonal block
                diag = ops.empty(ops.diag_size)
                if (self.p, self.q) == (pk, qk):
                    li, lj = self.rows_blk.index(k), self.cols_blk.index(k)
                    ops.factor_diag(work[li * b:(li + 1) * b, lj * b:(lj + 1) * b], diag, chain=True)
                self._bcast(diag, self._rank_of(pk, qk))
                self.diags.append(diag)
                if nbelow == 0:
                    self._pgrp.append(None)
                    break
                # 2. my rows of the panel (process column qk only), 3. replicate it: owner (ps, qk) -> slice ps
                Lall = ops.empty(nbelow * b, b)
                ops.wait(ev_wfree[k % 2])
                Wall = self._Wbuf[k % 2][:nbelow * b]
                if self.q == qk and mine:
                    li0, lj = self.rows_blk.index(mine[0]), self.cols_blk.index(k)
                    Bv = work[li0 * b:, lj * b:(lj + 1) * b]
                    Wloc = ops.panel(Bv, diag, chain=True)
                    Lall[offs[self.p]:offs[self.p + 1]].copy_(Bv)
                    Wall[offs[self.p]:offs[self.p + 1]].copy_(Wloc)
                for ps in range(P):
                    if groups[ps]:
                        self._bcast(Lall[offs[ps]:offs[ps + 1]], self._rank_of(ps, qk))
                        self._bcast(Wall[offs[ps]:offs[ps + 1]], self._rank_of(ps, qk))
                self._pgrp.append((Lall, groups))
                ev_panel = ops.record()
            # 4. trailing update of my blocks: A[I, J] -= W[I] L[J]^T for J > k, I >= J; block column k + 1 first
            ops.wait(ev_panel)
            ev_look = None
            mycols = [J for J in self.cols_blk if J > k]
            if mine and mycols:
                Wmine = Wall[offs[self.p]:offs[self.p + 1]]                      # my block rows, local order: contiguous
                pos = np.concatenate([np.arange(b) + offs[J % P] + groups[J % P].index(J) * b for J in mycols])
                Lmine = Lall.index_select(0, ops.index(pos))                       # L rows of my block columns
                li0, lj0 = self.rows_blk.index(mine[0]), self.cols_blk.index(mycols[0])
                if mycols[0] == k + 1:
                    ops.update_bc(work[li0 * b:, lj0 * b:(lj0 + 1) * b], Wmine, Lmine[:b], (P, Q), (self.p, self.q), li0, lj0)
                    ev_look = ops.record()
                    if len(mycols) > 1:
                        ops.update_bc(work[li0 * b:, (lj0 + 1) * b:], Wmine, Lmine[b:], (P, Q), (self.p, self.q), li0, lj0 + 1)
                else:
                    ops.update_bc(work[li0 * b:, lj0 * b:], Wmine, Lmine, (P, Q), (self.p, self.q), li0, lj0)
            if ev_look is None:
                ev_look = ops.record()
            ev_wfree[k % 2] = ops.record()
        ops.sync()
        self.inertia = tuple(int(v) for v in ops.counts_all(self.diags))   # one synchronisation at the very end
        self._solver = None
        return self.inertia

    def _global_panels(self):
        """L panels in global row order (what the replicated solver imports), from the per-owner layout of factor()"""
        out = []
        b = self.b
        for k, pg in enumerate(self._pgrp):
            if pg is None:
                out.append(None)
                continue
            Lall, groups = pg
            order = [I for gp in groups for I in gp]                 # grouped position -> global block
            inv = np.argsort(np.asarray(order))
            pos = np.concatenate([np.arange(b) + int(i) * b for i in inv])
            out.append(Lall.index_select(0, self.ops.index(pos)))
        return out

    def matvec(self, Xt):
        """Y = A X for nrhs x n device tensors (distributed original matrix: local product + all-reduce)."""
        Xc = Xt.index_select(1, self.ci).contiguous()
        if hasattr(self.ops, 'matvec_local'):
            part = self.ops.matvec_local(self.A0, Xc)                  # nrhs x (my rows), this repo's GEMV kernel
        else:
            part = torch.matmul(Xc, self.A0.t())                       # CPU reference backend (tests)
        Y = torch.zeros_like(Xt)
        Y[:, self.ri] = part
        if self.world > 1:
            dist.all_reduce(Y, group=self.group)
        return Y

    def solve_device(self, Bt, nrefine=1):
        """Bt: nrhs x n device tensor -> X (nrhs x n device tensor)."""
        if self._solver is None:
            if self._pgrp is not None:
                self.panels = self._global_panels()
            self._solver = self.ops.make_solver(self.n, self.diags, self.panels)
        X = self._solver(Bt.clone())
        for _ in range(nrefine):
            R = Bt - self.matvec(X)
            X = X + self._solver(R)
        return X

    def solve(self, B, nrefine=1):
        """B: n x nrhs (NumPy) -> X (NumPy).  Replicated triangular solves on the gathered factor + distributed
        iterative refinement."""
        Bt = self.ops.empty(B.shape[1], self.n)
        Bt.copy_(torch.from_numpy(np.ascontiguousarray(B.T)))
        return self.solve_device(Bt, nrefine).t().contiguous().cpu().numpy()


def choose_grid(world):
    # block-column-cyclic (1 x world): one broadcast per block column and a serial chain that never leaves a GPU; the 2-D
    # grids (2 x 2, 2 x 4) measured slower on NVSwitch, where every peer is one hop away at full bandwidth
    return (1, world)
