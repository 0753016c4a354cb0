"""Per-CTA timeline of the fused InfoNCE kernel (developer tool)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from passl_b200 import _lib, kernels as K
lib = _lib.load()
N, D, Kq, T = 256, 128, 65536, 0.2
q = torch.nn.functional.normalize(torch.randn(N, D, device="cuda"), dim=1)
k = torch.nn.functional.normalize(torch.randn(N, D, device="cuda"), dim=1)
queue = torch.nn.functional.normalize(torch.randn(Kq, D, device="cuda"), dim=1).bfloat16()
qb = q.bfloat16()
for _ in range(3):
    K.infonce_tc_fwd(qb, queue, pos=k, scale=1 / T)
dbg = torch.zeros(148 * 16, dtype=torch.int64, device="cuda")
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
flush.zero_()
torch.cuda.synchronize()
lib.passl_b200_infonce_tc_set_debug(dbg.data_ptr())
K.infonce_tc_fwd(qb, queue, pos=k, scale=1 / T)
torch.cuda.synchronize()
lib.passl_b200_infonce_tc_set_debug(None)
t = dbg.cpu().reshape(148, 16).double()
t0 = t[:, 0].min()
names = ["start", "after alloc+sync", "q_full", "q_ready arrive", "target done", "s_full t0", "t1", "t2", "t3", "t4", "t5", "t6", "t7", "loop end", "mma: q_ready"]
for i, n in enumerate(names):
    col = t[:, i]
    col = col[col > 0]
    if len(col):
        print("%-18s min %7.2f us  median %7.2f us  max %7.2f us" % (n, (col.min() - t0) / 1e3, (col.median() - t0) / 1e3, (col.max() - t0) / 1e3))
