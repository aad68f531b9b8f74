"""TEST INFRASTRUCTURE: CPU (torch/NumPy/SciPy) stand-in for pyipm_b200.dist_ldlt.CudaTileOps, so that the
block-cyclic indexing and the collective sequence of BlockCyclicLDLT can be exercised with gloo on CPU.
Same algebra (diagonal block  T = lu d lu',  L = B lu^-T d^-1,  W = L d,  C -= W L'), pivoting by
scipy.linalg.ldl inside the diagonal block."""
import numpy as np
import scipy.linalg
import torch


class RefTileOps(object):
    def __init__(self, block=8):
        self.b = block
        self.diag_size = 2 * block * block + 3

    def empty(self, *shape):
        return torch.empty(shape, dtype=torch.float64)

    def zeros(self, *shape):
        return torch.zeros(shape, dtype=torch.float64)

    # the look-ahead pipeline's stream plumbing degenerates to plain sequential execution on CPU
    def chain(self):
        import contextlib
        return contextlib.nullcontext()

    def record(self):
        return None

    def wait(self, ev):
        pass

    def sync(self):
        pass

    def factor_diag(self, Akk, diag, chain=False):
        b = self.b
        T = np.tril(Akk.numpy())
        T = T + np.tril(T, -1).T
        lu, d, _ = scipy.linalg.ldl(T, lower=True)
        w = np.linalg.eigvalsh(d)
        diag[:b * b] = torch.from_numpy(np.ascontiguousarray(lu).reshape(-1))
        diag[b * b:2 * b * b] = torch.from_numpy(np.ascontiguousarray(d).reshape(-1))
        diag[-3:] = torch.tensor([float(np.sum(w > 0)), float(np.sum(w < 0)), float(np.sum(w == 0))], dtype=torch.float64)

    def _unpack(self, diag):
        b = self.b
        return diag[:b * b].view(b, b).numpy(), diag[b * b:2 * b * b].view(b, b).numpy()

    def panel(self, Bblk, diag, chain=False, out=None):
        lu, d = self._unpack(diag)
        B = Bblk.numpy()
        W = np.linalg.solve(lu, B.T).T
        L = np.linalg.solve(d.T, W.T).T
        Bblk.copy_(torch.from_numpy(np.ascontiguousarray(L)))
        Wt = torch.from_numpy(np.ascontiguousarray(W))
        if out is not None:
            out.copy_(Wt)
            return out
        return Wt

    def update(self, Cv, W, L):
        Cv.sub_(W @ L.t())

    def update_bc(self, Cv, W, L, grid, coord, li0, lj0):
        b, (P, Q), (p, q) = self.b, grid, coord
        full = W @ L.t()
        for bi in range(Cv.shape[0] // b):
            for bj in range(Cv.shape[1] // b):
                if (li0 + bi) * P + p >= (lj0 + bj) * Q + q:
                    Cv[bi * b:(bi + 1) * b, bj * b:(bj + 1) * b] -= full[bi * b:(bi + 1) * b, bj * b:(bj + 1) * b]

    def counts_all(self, diags):
        tot = [0, 0, 0]
        for d in diags:
            for i, v in enumerate(d[-3:].tolist()):
                tot[i] += int(v)
        return tot

    def index(self, idx):
        return torch.from_numpy(idx)

    def make_solver(self, n, diags, panels):
        b = self.b
        Lf = np.zeros((n, n))
        Df = np.zeros((n, n))
        for k, diag in enumerate(diags):
            lu, d = self._unpack(diag)
            r0 = k * b
            Lf[r0:r0 + b, r0:r0 + b] = lu
            Df[r0:r0 + b, r0:r0 + b] = d
            if panels[k] is not None:
                Lf[r0 + b:, r0:r0 + b] = panels[k].numpy()

        def solve(Bt):
            y = np.linalg.solve(Lf, Bt.numpy().T)
            z = np.linalg.solve(Df, y)
            x = np.linalg.solve(Lf.T, z)
            return torch.from_numpy(np.ascontiguousarray(x.T))
        return solve
