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
    def _dup_mask(row_ids, col_ids, n_rows, n_cols, diag_shift):
        """True where row and column carry the same caption id off the positive pair."""
        m = row_ids.reshape(-1, 1) == col_ids.reshape(1, -1)
        rows = torch.arange(n_rows)
        dcol = rows - diag_shift
        has = (dcol >= 0) & (dcol < n_cols)
        m[rows[has], dcol[has]] = False
        return m

    @staticmethod
    def lse_stats_fused(x, y, scale, diag_shift, row_ids=None, col_ids=None):
        if row_ids is not None:
            c = x.double() @ y.double().T
            n_rows, n_cols = c.shape
            mask = ContractOps._dup_mask(row_ids, col_ids, n_rows, n_cols, diag_shift)
            rows = torch.arange(n_rows)
            dcol = rows - diag_shift
            has = (dcol >= 0) & (dcol < n_cols)
            diag = torch.zeros(n_rows, dtype=torch.float64)
            diag[has] = c[rows[has], dcol[has]]
            cm = c.masked_fill(mask, float("-inf"))
            row_max = cm.max(dim=1).values
            e = torch.exp(scale * (cm - row_max[:, None]))
            e[rows[has], dcol[has]] = 0.0
            col_ref = cm.max(dim=0).values + 0.25
            ec = torch.exp(scale * (cm - col_ref[None, :]))
            ec[rows[has], dcol[has]] = 0.0
            return row_max, e.sum(dim=1), diag, col_ref, ec.sum(dim=0)
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
             out_mul=None, out_dtype=None, row_ids=None, col_ids=None):
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
        if row_ids is not None:
            g[ContractOps._dup_mask(row_ids, col_ids, n_rows, n_cols, diag_shift)] = 0.0
        g[rows[has], dcol[has]] = -(w_row * xq[rows[has]] + w_col * yq[dcol[has]])
        g = g / (2.0 * n_global)
        dx = scale * (g @ y)
        if out_mul is not None:
            dx = dx * out_mul
        ds = (g * c).sum().reshape(1) if want_dscale else None
        return dx, ds

    @staticmethod
    def grad_both(x, y, x_stats, y_stats, scale, diag_shift, n_global, w_row, w_col, want_dscale,
                  out_mul=None, out_dtypes=None, window=None, row_ids=None, col_ids=None):
        """Contract of vlpclip_grad_both: (dX, dY, dscale) of ONE sweep; dY = G^T X is the other
        direction's dX with the roles of X and Y (and of the weights) swapped."""
        dx, ds = ContractOps.grad(x, y, x_stats, y_stats, scale, diag_shift, n_global, w_row, w_col,
                                  want_dscale, out_mul, None, row_ids, col_ids)
        dy, _ = ContractOps.grad(y, x, y_stats, x_stats, scale, -diag_shift, n_global, w_col, w_row,
                                 False, out_mul, None, col_ids, row_ids)
        return dx, dy, ds

    @staticmethod
    def to_backward_operand(x):
        return x


class _EmulatedWindow:
    """What functional.PeerWindow is to the plan: per-parity reduce-scatter slots and gathered
    operand; 'peer stores' are emulated with gloo collectives at the moment the kernel would run."""

    def __init__(self, group, world, rank, rows, d):
        self.group, self.world, self.rank, self.rows, self.d = group, world, rank, rows, d
        self.slots = {}
        self.parity = 0
        self.log = []


class WindowContractOps(ContractOps):
    """ContractOps plus the optional peer-window ops of the sharded plan (push_gather,
    peer_window / grad_scatter / scatter_finish), so that the fused-collective branches of
    sharded.forward_plan / backward_plan run on CPU under gloo."""
    windows = {}

    @classmethod
    def peer_window(cls, group, world, rank, rows, d, device):
        key = (id(group), rows, d)
        if key not in cls.windows:
            cls.windows[key] = _EmulatedWindow(group, world, rank, rows, d)
        return cls.windows[key]

    @classmethod
    def push_gather(cls, i_loc, t_loc, group, world, rank):
        import torch.distributed as dist
        w = cls.peer_window(group, world, rank, t_loc.shape[0], t_loc.shape[1], t_loc.device)
        t_all = torch.empty(world * t_loc.shape[0], t_loc.shape[1], dtype=t_loc.dtype)
        dist.all_gather_into_tensor(t_all, t_loc.contiguous(), group=group)   # = the peer stores
        dist.all_reduce(torch.zeros(1), group=group)                          # closing collective
        w.log.append("push_gather")
        return i_loc, t_all

    @classmethod
    def grad_scatter(cls, x, y, x_stats, y_stats, scale, diag_shift, n_global, w_row, w_col,
                     want_dscale, window):
        import torch.distributed as dist
        dx, ds = ContractOps.grad(x, y, x_stats, y_stats, scale, diag_shift, n_global, w_row, w_col,
                                  want_dscale)
        window.parity ^= 1
        # rank s stores rows [o*rows, (o+1)*rows) of its partial into slot s of owner o
        parts = [torch.empty_like(dx) for _ in range(window.world)]
        dist.all_gather(parts, dx.contiguous(), group=window.group)
        lo = window.rank * window.rows
        window.slots[window.parity] = [p[lo:lo + window.rows].clone() for p in parts]
        window.log.append("grad_scatter")
        return window.parity, ds

    @classmethod
    def grad_both(cls, x, y, x_stats, y_stats, scale, diag_shift, n_global, w_row, w_col, want_dscale,
                  out_mul=None, out_dtypes=None, window=None, row_ids=None, col_ids=None):
        import torch.distributed as dist
        if window is None:
            return ContractOps.grad_both(x, y, x_stats, y_stats, scale, diag_shift, n_global, w_row,
                                         w_col, want_dscale, out_mul, None, None, row_ids, col_ids)
        dx, ds = ContractOps.grad(x, y, x_stats, y_stats, scale, diag_shift, n_global, w_row, w_col,
                                  want_dscale, out_mul, None, row_ids, col_ids)
        # the dY rows leave unscaled through the owners' slots (out_mul is applied by scatter_finish)
        dy, _ = ContractOps.grad(y, x, y_stats, x_stats, scale, -diag_shift, n_global, w_col, w_row, False,
                                 None, None, col_ids, row_ids)
        window.parity ^= 1
        parts = [torch.empty_like(dy) for _ in range(window.world)]
        dist.all_gather(parts, dy.contiguous(), group=window.group)
        lo = window.rank * window.rows
        window.slots[window.parity] = [p[lo:lo + window.rows].clone() for p in parts]
        window.log.append("grad_both")
        return dx, window.parity, ds

    @classmethod
    def scatter_finish(cls, window, parity, out_mul, out_dtype, device):
        acc = window.slots[parity][0].clone()
        for s in range(1, window.world):
            acc += window.slots[parity][s]
        window.log.append("scatter_finish")
        return acc * out_mul if out_mul is not None else acc
