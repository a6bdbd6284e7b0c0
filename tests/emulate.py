"""CPU emulation of the *addressing semantics* of b200_conv2d_fwd (include/b200diff.h): tap tables, parity
planes, weight packing and the 4-phase nearest-2x decomposition.  Used by the CPU test-suite to validate the host
logic that feeds the CUDA kernel; it is test infrastructure, never a product path."""
import torch


def conv_emulate(a0, w_packed, N, B, Ho, Wo, taps0, a1=None, tap1=(0, 0, 0), w_rows_per_phase=None):
    """a0: float [B, planes, Hs, Ws, C]; w_packed: float [rows, K]; returns float [B, out_H, out_W, N]."""
    phases = len(taps0)
    up = 2 if phases == 4 else 1
    out = torch.zeros(B, Ho * up, Wo * up, N, dtype=torch.float32)
    rpp = w_rows_per_phase if w_rows_per_phase is not None else N

    def gather(src, dw, dh, pl):
        _, _, Hs, Ws, C = src.shape
        g = torch.zeros(B, Ho, Wo, C)
        ys = torch.arange(Ho) + dh
        xs = torch.arange(Wo) + dw
        vy = (ys >= 0) & (ys < Hs)
        vx = (xs >= 0) & (xs < Ws)
        sub = src[:, pl][:, ys[vy]][:, :, xs[vx]]
        gy = torch.nonzero(vy).flatten()
        gx = torch.nonzero(vx).flatten()
        g[:, gy[:, None], gx[None, :]] = sub
        return g

    for ph in range(phases):
        cols = [gather(a0, dw, dh, pl) for (dw, dh, pl) in taps0[ph]]
        if a1 is not None:
            cols.append(gather(a1, *tap1))
        A = torch.cat(cols, dim=-1)  # [B, Ho, Wo, K]
        Wm = w_packed[ph * rpp: ph * rpp + N].float()
        res = A @ Wm.t()
        out[:, (ph >> 1)::up, (ph & 1)::up] = res
    return out
