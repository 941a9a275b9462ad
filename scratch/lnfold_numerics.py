"""Numerics of folding LayerNorm into the consuming GEMM algebraically: fp16(LN(x)) @ fp16(W)^T   vs
   rstd * (fp16(x) @ fp16(gamma*W)^T - mean * s) + c   against the fp32 result, on the real residual streams."""
import sys, torch
sys.path.insert(0,'/root/repo')
import torch.nn.functional as F
from oracle import ref_model as R
from oracle.synth import synth_images, synth_state_dict
from oracle.arch import ModelConfig
cfg=ModelConfig(); sd=synth_state_dict(cfg, seed=0)
torch.set_num_threads(16)
imgs=synth_images(2, seed=1234)
p="encoder.features."
with torch.no_grad():
    outs=R.encoder_stages(imgs, sd)
    # residual stream entering each block = previous entry in outs (patch embed / block / merge)
    idx=0; x=outs[0]
    res=[]
    for s in range(4):
        fi=1+2*s
        for j in range(R.DEPTHS[s]):
            bp=f"{p}{fi}.{j}."
            for norm, lin in (("norm1","attn.qkv"),("norm2","mlp.0")):
                if norm=="norm2":
                    xin = x + R.window_attention(R._ln(x, sd, bp+"norm1"), sd, bp, R.HEADS[s], 0 if j%2==0 else R.WINDOW//2)
                else:
                    xin = x
                g, b = sd[bp+norm+".weight"], sd[bp+norm+".bias"]
                W, bias = sd[bp+lin+".weight"], sd[bp+lin+".bias"]
                X = xin.reshape(-1, xin.shape[-1])
                exact = F.linear(F.layer_norm(X,(X.shape[-1],),g,b,1e-5).double(), W.double(), bias.double())
                cur = F.linear(F.layer_norm(X,(X.shape[-1],),g,b,1e-5).half().double(), W.half().double(), bias.double())
                mean = X.double().mean(-1, keepdim=True); var = X.double().var(-1, unbiased=False, keepdim=True); rstd = (var+1e-5).rsqrt()
                Wg = (W*g).half().double(); sN = Wg.sum(-1); cN = (W.double()@b.double()) + bias.double()
                alg = rstd*(X.half().double()@Wg.t() - mean*sN) + cN
                # variant: centre x by the row mean BEFORE rounding (needs the mean at production time) 
                e_cur=(cur-exact).abs().max().item(); e_alg=(alg-exact).abs().max().item()
                res.append((bp+lin, X.abs().max().item(), (mean.abs()/ (var.sqrt()+1e-9)).max().item(), e_cur, e_alg, exact.abs().max().item()))
            x = outs[idx+1]; idx+=1
        if s<3:
            x = outs[idx+1]; idx+=1
for r in res: print("%-38s |x|max %7.2f  max|mean|/std %6.2f  err fp16(LN) %.2e  err algebraic %.2e  (|out|max %.1f)"%r)
