"""SelfAttention_Layer — drop-in for `models/submodules.py:178-238` of the reference ("next" row 1 of SURVEY.md
§8f: the attention front-end of the AutoInt / STAR / MMOE 'usetrans' baselines), run through the same per-sample
field-attention kernel code as SATrans.  Same constructor, same state-dict keys (W_Query, W_Key, W_Value, W_Out,
W_Res, layer_norm.*).  Forward only (inference, and train-mode dropout with the counter-based RNG); the backward of
this layer is not built."""
import torch
import torch.nn as nn

from . import _lib


class SelfAttention_Layer(nn.Module):
    def __init__(self, embedding_size, head_num=2, use_res=True, scaling=True, seed=1024, device='cpu'):
        super().__init__()
        if head_num <= 0:
            raise ValueError('head_num must be a int > 0')
        if embedding_size % head_num != 0:
            raise ValueError('embedding_size is not an integer multiple of head_num!')
        if embedding_size != _lib.SAT_D or head_num not in (1, 2, 4, 8):
            raise NotImplementedError("kernels are built for embedding_size=32 and head_num in {1,2,4,8}")
        if not scaling:
            raise NotImplementedError("scaling=False is not built (every call site of the reference uses the default)")
        self.att_embedding_size = embedding_size // head_num
        self.head_num, self.use_res, self.scaling, self.seed = head_num, use_res, scaling, seed
        d = embedding_size
        self.W_Query = nn.Parameter(torch.empty(d, d))
        self.W_Key = nn.Parameter(torch.empty(d, d))
        self.W_Value = nn.Parameter(torch.empty(d, d))
        self.W_Out = nn.Parameter(torch.empty(d, d))          # declared, never used (reference submodules.py:195)
        self.layer_norm = nn.LayerNorm(d, eps=1e-6)
        if use_res:
            self.W_Res = nn.Parameter(torch.empty(d, d))
        for t in self.parameters():                            # reference submodules.py:203-204
            nn.init.normal_(t, mean=0.0, std=0.05)
        self.dropout_rate = 0.1
        self._calls = 0
        self.to(device)

    def forward(self, inputs):
        if len(inputs.shape) != 3:
            raise ValueError("Unexpected inputs dimensions %d, expect to be 3 dimensions" % (len(inputs.shape)))
        if not inputs.is_cuda:
            raise _lib.SatError('SelfAttention_Layer needs CUDA tensors; there is no CPU fallback')
        x = inputs.detach().contiguous().float()
        B, F, d = x.shape
        out = torch.empty_like(x)
        self._calls += 1
        ptr = lambda t: t.detach().contiguous().data_ptr()
        _lib.check(_lib.lib().sat_selfatt_fwd(
            x.data_ptr(), out.data_ptr(), ptr(self.W_Query), ptr(self.W_Key), ptr(self.W_Value),
            ptr(self.W_Res) if self.use_res else None, ptr(self.layer_norm.weight), ptr(self.layer_norm.bias),
            B, F, self.head_num, int(self.use_res), int(self.training), float(self.dropout_rate),
            (int(torch.initial_seed()) + self._calls) & (2 ** 64 - 1), torch.cuda.current_stream().cuda_stream),
            'sat_selfatt_fwd')
        return out
