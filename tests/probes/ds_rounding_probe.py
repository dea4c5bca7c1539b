"""Probe behind the q/k gradient exception of tests/test_model_parity_gpu.py: the oracle (fp32, CPU) with ONLY the
attention backward inputs / the dS and P tiles rounded (bf16, scaled fp16) -> worst relative error of the q/k gradients.
Build-container tool (imports oracle/); prints three lines."""
import sys, math, torch
import os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from oracle import codon_gpt_oracle as O
torch.manual_seed(0)
# C2-like: 6L4H d256 rope swiglu, T=512, B=2
ctor=dict(vocab_size=68, block_size=512, n_layer=6, n_head=4, n_embd=256, dropout=0.0, label_smoothing=0.05, use_sdpa=True, use_rope=True, use_swiglu=True)
cfg=O.make_cfg(**ctor); sd=O.init_state_dict(cfg, seed=1337, emb_scale=0.02)
idx,tgt=O.synthetic_batch(2,512,seed=1337,realistic=True)
MODE={'m':'exact'}
class Attn(torch.autograd.Function):
    @staticmethod
    def forward(ctx,q,k,v,mask):
        hd=q.shape[-1]
        s=(q@k.transpose(-2,-1))/math.sqrt(hd)
        s=s.masked_fill(~mask,float('-inf'))
        p=torch.softmax(s,-1)
        ctx.save_for_backward(q,k,v,p); 
        return p@v
    @staticmethod
    def backward(ctx,do):
        q,k,v,p=ctx.saved_tensors; hd=q.shape[-1]; sc=1/math.sqrt(hd)
        r=lambda t: t
        m=MODE['m']
        if m!='exact':
            rb=lambda t: t.to(torch.bfloat16).float()
            q_,k_,v_,do_=rb(q),rb(k),rb(v),rb(do)
        else: q_,k_,v_,do_=q,k,v,do
        dp=do_@v_.transpose(-2,-1)
        delta=(p*dp).sum(-1,keepdim=True)
        ds=p*(dp-delta)
        pb=p
        if m=='bf16':
            ds=ds.to(torch.bfloat16).float(); pb=p.to(torch.bfloat16).float()
        elif m=='fp16':
            scale=2.0**(14-math.ceil(math.log2(ds.abs().max().item())))
            ds=(ds*scale).to(torch.float16).float()/scale; pb=p.to(torch.bfloat16).float()
        elif m=='inputs_only':
            pass
        dv=pb.transpose(-2,-1)@do_
        dq=(ds@k_)*sc; dk=(ds.transpose(-2,-1)@q_)*sc
        return dq,dk,dv,None
# monkeypatch oracle attention
import types
src=O._attention
def _attention(sd, pre, x, cfg, mask_bool, cos, sin):
    B,T,d=x.shape; H=cfg['n_head']; Hk=cfg.get('n_kv_head') or H; hd=d//H
    q=O.linear(x,sd[pre+'query.weight'],sd[pre+'query.bias']).view(B,T,H,hd).transpose(1,2)
    k=O.linear(x,sd[pre+'key.weight'],sd[pre+'key.bias']).view(B,T,Hk,hd).transpose(1,2)
    v=O.linear(x,sd[pre+'value.weight'],sd[pre+'value.bias']).view(B,T,Hk,hd).transpose(1,2)
    if cos is not None:
        q=O.apply_rope(q,cos,sin); k=O.apply_rope(k,cos,sin)
    y=Attn.apply(q,k,v,mask_bool.expand(B,H,T,T))
    y=y.transpose(1,2).reshape(B,T,d)
    return O.linear(y,sd[pre+'proj.weight'],sd[pre+'proj.bias']), None
import inspect
print(inspect.signature(src))
O._attention=_attention
res={}
for mode in ['exact','inputs_only','bf16','fp16']:
    MODE['m']=mode
    total,parts,out,grads=O.loss_and_grads(sd,cfg,idx,tgt)
    res[mode]=grads
ref=res['exact']
for mode in ['inputs_only','bf16','fp16']:
    worst=0;name=None
    for kname,g in ref.items():
        if '.attn.query.' in kname or '.attn.key.' in kname:
            den=g.norm().item()
            if den<1e-12: continue
            e=((res[mode][kname]-g).norm()/den).item()
            if e>worst: worst,name=e,kname
    print(mode,'worst q/k rel err',worst,name)
