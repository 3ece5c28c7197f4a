import torch, json, sys
B,H,W=8,352,1216
dev=torch.device("cuda:0")
x64=torch.randn(B,64,H,W,device=dev)
def timed(fn,n=5):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return round(e0.elapsed_time(e1)/n,3)
out={}
for n in (1,24,26):
    g=torch.randn(B,n,H,W,device=dev); w=torch.randn(n,64,3,3,device=dev)
    out["dgrad_%d_to_64"%n]=timed(lambda: torch.nn.grad.conv2d_input(x64.shape,w,g,stride=1,padding=1))
    out["wgrad_%d_x_64"%n]=timed(lambda: torch.nn.grad.conv2d_weight(x64,w.shape,g,stride=1,padding=1))
    gcl=g.contiguous(memory_format=torch.channels_last); xcl=x64.contiguous(memory_format=torch.channels_last); wcl=w.contiguous(memory_format=torch.channels_last)
    out["dgrad_%d_to_64_cl"%n]=timed(lambda: torch.nn.grad.conv2d_input(xcl.shape,wcl,gcl,stride=1,padding=1))
    out["wgrad_%d_x_64_cl"%n]=timed(lambda: torch.nn.grad.conv2d_weight(xcl,wcl.shape,gcl,stride=1,padding=1))
gs=[torch.randn(B,n,H,W,device=dev) for n in (1,24,1)]
out["cat_g"]=timed(lambda: torch.cat(gs,1))
out["bias_sums"]=timed(lambda: [g.sum(dim=(0,2,3)) for g in gs])
out["relu_sig_masks"]=timed(lambda: (gs[0]*(gs[0]>0).float(), gs[2]*gs[2]*(1-gs[2])))
print(json.dumps(out))
