import sys, torch
sys.path.insert(0, '/root/repo')
from nlspn_eccv20_b200 import heads
dev = torch.device('cuda:0')
B,H,W,K = [int(a) for a in sys.argv[1:5]]
g = torch.Generator().manual_seed(1)
N3 = 3*(K*K-1)
x = [torch.randn(B,64,H,W,generator=g).to(dev) for _ in range(4)]
w = [(0.03*torch.randn(n,128,3,3,generator=g)).to(dev) for n in (1,N3,1)]
b = [(0.1*torch.randn(n,generator=g)).to(dev) for n in (1,N3,1)]
o = heads.fused_heads(x[0],x[1],x[2],x[3],w[0],b[0],w[1],b[1],w[2],b[2],K)
torch.cuda.synchronize()
r = heads.reference_heads(x[0],x[1],x[2],x[3],w[0],b[0],w[1],b[1],w[2],b[2])
print([float((a-c).abs().max()) for a,c in zip(o,r)])
