"""Test-only numpy/torch restatement of the KERNEL CONTRACTS used by ``vlp_b200.sharded``
(``ops`` interface) so that the collective orchestration can run on CPU tensors under gloo.
Mirrors what lse_fwd.cu / grad_bwd.cu compute per call (fp64, no tiling)."""
import math

import torch


class ContractOps:
    @staticmethod
    def lse_stats(x, y, scale, diag_shift):
        x = x.double(); y = y.double()
        c = x @ y.T
        n_rows, n_cols = c.shape
        rows = torch.arange(n_rows)
        dcol = rows - diag_shift
        has = (dcol >= 0) & (dcol < n_cols)
        diag = torch.zeros(n_rows, dtype=torch.float64)
        diag[has] = c[rows[has], dcol[has]]
        row_max = c.max(dim=1).values
        e = torch.exp(scale * (c - row_max[:, None]))
        e[rows[has], dcol[has]] = 0.0                     # positive pair left out of the sum
        return row_max, e.sum(dim=1), diag

    @staticmethod
    def lse_stats_fused(x, y, scale, diag_shift):
        row_max, row_l, diag = ContractOps.lse_stats(x, y, scale, diag_shift)
        c = x.double() @ y.double().T
        n_rows, n_cols = c.shape
        rows = torch.arange(n_rows)
        dcol = rows - diag_shift
        has = (dcol >= 0) & (dcol < n_cols)
        col_ref = c.max(dim=0).values + 0.25            # any upper reference is allowed
        e = torch.exp(scale * (c - col_ref[None, :]))
        e[rows[has], dcol[has]] = 0.0
        return row_max, row_l, diag, col_ref, e.sum(dim=0)

    @staticmethod
    def merge_stats(part_max, part_l, diag, scale):
        if part_max.dim() == 1:
            part_max = part_max[None]; part_l = part_l[None]
        m = part_max.max(dim=0).values
        l = (part_l * torch.exp(scale * (part_max - m[None]))).sum(dim=0)
        t = torch.exp(scale * (diag - m))
        tot = l + t
        lg2l = torch.log2(tot)                              # log2(sum) - k*max
        q = l / tot
        row_loss = torch.log(tot) + scale * (m - diag)
        return m, lg2l, q, row_loss

    @staticmethod
    def loss_sums(row_loss, col_loss):
        return torch.stack([row_loss.sum(), col_loss.sum()])

    @staticmethod
    def grad(x, y, x_stats, y_stats, scale, diag_shift, n_global, w_row, w_col, want_dscale,
             out_mul=None, out_dtype=None):
        x = x.double(); y = y.double()
        xm, xlg, xq = x_stats
        ym, ylg, yq = y_stats
        k = scale * math.log2(math.e)
        c = x @ y.T
        p_row = torch.exp2(k * (c - xm[:, None]) - xlg[:, None])
        p_col = torch.exp2(k * (c - ym[None, :]) - ylg[None, :])
        g = w_row * p_row + w_col * p_col
        n_rows, n_cols = c.shape
        rows = torch.arange(n_rows)
        dcol = rows - diag_shift
        has = (dcol >= 0) & (dcol < n_cols)
        g[rows[has], dcol[has]] = -(w_row * xq[rows[has]] + w_col * yq[dcol[has]])
        g = g / (2.0 * n_global)
        dx = scale * (g @ y)
        if out_mul is not None:
            dx = dx * out_mul
        ds = (g * c).sum().reshape(1) if want_dscale else None
        return dx, ds

    @staticmethod
    def to_backward_operand(x):
        return x
