"""Debug probe: tensor-core weight-gradient GEMM (K-major transposed staging, SS MMA) vs torch."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from satrans_b200 import _lib
torch.manual_seed(0)
A = torch.randn(128, 32, device='cuda'); G = torch.randn(128, 32, device='cuda')
out = torch.zeros(32, 32, device='cuda'); raw = torch.zeros(128, 32, device='cuda')
_lib.check(_lib.lib().sat_debug_wgrad(A.data_ptr(), G.data_ptr(), out.data_ptr(), raw.data_ptr(), None), 'wgrad')
torch.cuda.synchronize()
ref = A.double().t() @ G.double()
print('max |out|', out.abs().max().item(), 'max |ref|', ref.abs().max().item())
print('err A^T G  ', (out.double() - ref).abs().max().item(), ' rel', ((out.double() - ref).abs().max() / ref.abs().max()).item())
print('rows >= 32 finite:', bool(torch.isfinite(raw).all()))
