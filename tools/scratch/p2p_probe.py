"""Debug probe (torchrun, 2 GPUs): which kinds of access to CUDA-IPC peer memory work."""
import os, sys, torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from satrans_b200.dist import DistContext
from satrans_b200 import _lib
rank, local = int(os.environ['RANK']), int(os.environ['LOCAL_RANK'])
torch.cuda.set_device(local)
dist.init_process_group('nccl', device_id=torch.device('cuda:%d' % local))
ctx = DistContext()
for n in (16, 1 << 20):
    t = torch.zeros(n, dtype=torch.float32, device='cuda')
    ptrs = ctx.exchange_pointers(t)
    out = torch.zeros(1, device='cuda')
    ctx.barrier('cuda'); torch.cuda.synchronize()
    for mode in (4, 0, 1, 2, 3):
        try:
            _lib.check(_lib.lib().sat_debug_poke(ptrs[1 - rank], mode, out.data_ptr(), None), 'poke')
            torch.cuda.synchronize()
            print(rank, 'n', n, 'mode', mode, 'ok', flush=True)
        except Exception as e:
            print(rank, 'n', n, 'mode', mode, 'FAILED', str(e)[:80], flush=True)
            break
    ctx.barrier('cuda'); torch.cuda.synchronize()
    print(rank, 'own', t[:4].tolist(), flush=True)
dist.destroy_process_group()
