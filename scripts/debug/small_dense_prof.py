"""One launch each of the small dense kernels at the C2 block-stage sizes (for ncu captures): batched DMMA Gram, Jacobi
(float64 n = 50, float32 n = 60), block CholQR2, the symmetric float64 product."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from localmd_b200 import ops

torch.manual_seed(0)
nb = 2601
x = torch.randn(nb, 50, 5000, device="cuda") * torch.logspace(0, -2, 50, device="cuda")[None, :, None]
g = ops.gram_rows(x)
ops.jacobi_eigh(g.clone(), mode=0)
x6 = torch.randn(nb, 60, 500, device="cuda") * torch.logspace(0, -2, 60, device="cuda")[None, :, None]
ops.jacobi_eigh(ops.gram_rows(x6), mode=0, sweeps_f32=True)
s = torch.randn(nb, 400, 52, device="cuda")
s[:, :, 50:] = 0
ops.block_orth(s, 50)
a = torch.randn(1650, 20000, device="cuda")
ops.sym_product_f64(a)
torch.cuda.synchronize()
print("done")
