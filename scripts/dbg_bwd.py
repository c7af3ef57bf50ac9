import sys, torch
sys.path.insert(0,'.')
from generative_recommenders_b200 import _lib
from generative_recommenders_b200.common import HammerKernel, generate_sparse_seq_len
from generative_recommenders_b200.ops.hstu_attention import hstu_mha
dev=torch.device('cuda')
torch.manual_seed(3)
B,H,d,lmax=int(sys.argv[1]),int(sys.argv[3]) if len(sys.argv)>3 else 2,32,int(sys.argv[2])
lengths=generate_sparse_seq_len(B,lmax,0.95,dev)
off=torch.zeros(B+1,dtype=torch.int64,device=dev); off[1:]=torch.cumsum(lengths,0)
L=int(off[-1])
x=torch.empty(L,H,3*d,device=dev,dtype=torch.bfloat16).uniform_(-0.5,0.5)
q,k,v=torch.split(x,[d,d,d],dim=-1)
res={}
for impl in (_lib.IMPL_GENERIC,_lib.IMPL_AUTO):
    qq,kk,vv=(t.detach().clone().requires_grad_() for t in (q,k,v))
    o=hstu_mha(lmax,0.17,qq,kk,vv,off,kernel=HammerKernel.CUDA,impl=impl)
    o.backward(torch.ones_like(o))
    torch.cuda.synchronize()
    res[impl]=(qq.grad.float(),kk.grad.float(),vv.grad.float())
for n,a,b in zip('qkv',res[_lib.IMPL_AUTO],res[_lib.IMPL_GENERIC]):
    print(n,'relerr',float((a-b).norm()/b.norm()))
