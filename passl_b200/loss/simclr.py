"""SimCLR NT-Xent + CO2 as one autograd node (simclr_contrastive_head.py:42-102), optionally with all-gathered negatives
(the BASELINE "global bs=4096, all-gather negatives" variant; the reference head is local-only and carries an unused
`multi_rank` flag, :35-40).

   R = [h1; h2] (2n x d, local);  Z = all_gather(R) (rank-interleaved [h1_r; h2_r] blocks, 2m x d)
   S  = R . Z^T / T                                tcgen05 GEMM, fp32 out (stays in L2)
   loss, dS from the row-pair kernel (ntxent.cu)
   dR = dS . Z / T ,  dZ = dS^T . R / T            two more tcgen05 GEMMs; dZ flows back through the differentiable
                                                   all-gather (reduce-scatter) when world > 1.
"""
import torch

from .. import kernels as K


class _NTXentCO2(torch.autograd.Function):
    @staticmethod
    def forward(ctx, R, Z, n, m, rank, temperature, co2_weight):
        Rb = K.cast_bf16(R.contiguous())
        Zb = Rb if Z is None else K.cast_bf16(Z.contiguous())
        S = K.gemm(Rb, Zb, out_dtype=torch.float32, alpha=1.0 / temperature)
        out, ws = K.ntxent_co2_fwd(S, n, m, rank, co2_weight)
        ctx.saved = (Rb, Zb, S, ws)
        ctx.cfg = (n, m, rank, temperature, co2_weight, Z is None)
        ctx.mark_non_differentiable(out[1])
        return out[0], out[1]

    @staticmethod
    def backward(ctx, dloss, _acc):
        Rb, Zb, S, ws = ctx.saved
        n, m, rank, T, w, local = ctx.cfg
        dS = K.ntxent_co2_bwd(S, ws, n, m, rank, w, dloss=dloss.contiguous().float())
        dR = K.gemm(dS, Zb, b_t=True, out_dtype=torch.float32, alpha=1.0 / T)            # [2n, d]  (row side)
        ctx.saved = None
        if local:    # Z is R itself: add the column-side gradient in the same buffer (atomic accumulate epilogue)
            K.gemm(dS, Rb, a_t=True, b_t=True, out=dR, accumulate=True, alpha=1.0 / T)
            return dR, None, None, None, None, None, None
        dZ = K.gemm(dS, Rb, a_t=True, b_t=True, out_dtype=torch.float32, alpha=1.0 / T)  # [2m, d]  (column side)
        return dR, dZ, None, None, None, None, None


def ntxent_co2(con, n, temperature, co2_weight=3.0, gather=None, rank=0):
    """con: fp32 [2n, d] = [h1; h2] normalised.  gather: differentiable all-gather (None -> local negatives only).
    Returns (loss, acc1)."""
    if gather is None:
        return _NTXentCO2.apply(con, None, n, n, 0, temperature, co2_weight)
    Z = gather(con)
    return _NTXentCO2.apply(con, Z, n, Z.shape[0] // 2, rank, temperature, co2_weight)
