"""Runs tools/probe/mn_probe.cu (MN-major SWIZZLE_128B tf32 operands) on the GPU and reports which staging / descriptor variant
reproduces A^T G.  Build first: make -C satrans_b200/csrc probe"""
import ctypes as C, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
lib = C.CDLL(os.path.join(ROOT, 'satrans_b200', 'lib', 'libsat_probe.so'))
lib.sat_mn_probe.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p]
torch.manual_seed(0)
def tf32(x): return (x.view(torch.int32) & -8192).view(torch.float32)
A = tf32(torch.randn(128, 128)).cuda(); G = tf32(torch.randn(128, 32)).cuda()
ref = (A.double().t() @ G.double()).float()
for flags in (4 | (1 << 4), 4 | (2 << 4), 1 | (2 << 4)):
    out = torch.full((128, 32), float('nan'), device='cuda')
    rc = lib.sat_mn_probe(A.data_ptr(), G.data_ptr(), out.data_ptr(), flags, None)
    try:
        torch.cuda.synchronize()
        err = (out - ref).abs().max().item()
        print('flags 0x%x: rc %d max err %.3e  (ref max %.2f, out absmax %.2f)' % (flags, rc, err, ref.abs().max().item(), out.abs().max().item()), flush=True)
    except Exception as e:
        print('flags 0x%x: CUDA error %s' % (flags, e)); sys.exit(1)
