"""Stand-in for flowtorch.analysis.SVD: only .s/.U/.V/.rank are read by the reference (utils.py:330-346)."""
import torch as pt


class SVD:
    def __init__(self, data_matrix, rank=None):
        u, s, vh = pt.linalg.svd(data_matrix, full_matrices=False)
        r = s.numel() if rank is None else min(int(rank), s.numel())
        self.rank, self.U, self.s, self.V = r, u[:, :r], s[:r], vh.conj().T[:, :r]
