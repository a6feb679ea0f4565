"""TEST INFRASTRUCTURE (CPU only, uses the oracle).  Numerical study behind DESIGN.md 5.3 / 10: what happens to the network output
when the relative-position bias table 16 * sigmoid(CPB_MLP(.)) (mixed_attn_block_efficient.py:41-47) is rounded to a 16-bit format
before it is added to the scores -- the step that would halve the bias traffic of the attention kernel's softmax warps.
fp16c = the table minus 8 (softmax is invariant to a per-head constant) rounded to fp16.

    python oracle/study_bias_fp16.py

Measured in the build container (64x64 input, fp32 everywhere else, PSNR(candidate, unrounded) / |dPSNR vs a random GT| / max-abs):
    base  init    fp16 107.0 dB 9.5e-07 2.3e-05 | fp16c 125.4 dB 0.0e+00 2.7e-06 | bf16 86.8 dB 7.7e-05 2.0e-04
    base  spread  fp16  36.2 dB 9.4e-03 7.1e-02 | fp16c  37.6 dB 7.3e-03 5.9e-02 | bf16 18.2 dB 3.2e-02 6.3e-01
    small spread  fp16  56.2 dB 3.4e-04 7.7e-03 | fp16c  61.6 dB 9.0e-05 4.4e-03 | bf16 37.9 dB 2.4e-03 7.4e-02
("spread" = the stress weights of oracle.synth_state_dict, a deliberately ill-conditioned network that amplifies ANY rounding;
"init" = weights distributed like the reference constructor's, the style every gate in tests/ and bench.py uses.)
"""
import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, HERE)
import grl_oracle as orc
from _pkgload import load_package
pkg=load_package()
torch.set_num_threads(8)
real_sig=torch.sigmoid
mode=[None]
def sig(x):
    y=real_sig(x)
    if mode[0]=='fp16': return (16*y).half().float()/16
    if mode[0]=='fp16c': return ((16*y-8).half().float()+8)/16
    if mode[0]=='bf16': return (16*y).bfloat16().float()/16
    return y
for variant,task,scale,size,style in (("base","sr",4,64,"init"),("base","sr",4,64,"spread"),("small","sr",4,64,"spread")):
    cfg=pkg.configs.grl_config(variant,task,scale,size)
    sd=orc.synth_state_dict(cfg,seed=0,style=style)
    x=orc.synth_input((1,3,size,size),seed=1234)
    gt=torch.rand(1,3,size*scale,size*scale,generator=torch.Generator().manual_seed(9))
    outs={}
    for m in (None,'fp16','fp16c','bf16'):
        mode[0]=m
        torch.sigmoid=sig
        with torch.no_grad(): outs[m]=orc.grl_forward(sd,cfg,x)
        torch.sigmoid=real_sig
    ref=outs[None]
    for m in ('fp16','fp16c','bf16'):
        p=(-10*torch.log10(((outs[m]-ref)**2).mean())).item()
        d=abs(orc.psnr(outs[m],gt,scale).mean().item()-orc.psnr(ref,gt,scale).mean().item())
        print(variant,style,m,f"PSNR(cand,ref)={p:.1f} dB  dPSNR_vs_GT={d:.2e} max-abs={(outs[m]-ref).abs().max().item():.2e}")
