"""Eager vs CUDA-graph step time for the C3 workload (diagnostic)."""
import os, sys, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "genomics-lm_b200"))
import bench
from codonlm_b200 import TinyGPT
from codonlm_b200.trainer import TrainStep
L = int(os.environ.get("LAYERS", 12)); B = int(os.environ.get("BATCH", 64)); T = 1024
torch.manual_seed(1337)
model = TinyGPT(**bench.workload_ctor(L, T))
with torch.no_grad():
    model.tok_emb.weight.mul_(0.02); model.pos_emb.weight.mul_(0.02)
model = model.cuda().train()
step = TrainStep(model, lr=3e-4, lr_embedding=3e-4, weight_decay=0.05, offset_weights=bench.OFFSET_W, termination_loss_weight=bench.TERM_W)
x, y = bench.synthetic_tokens(B, T, 1337); x, y = x.cuda(), y.cuda()
def timeit(n=5):
    torch.cuda.synchronize(); s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(n): loss = step.step(x, y)
    e.record(); torch.cuda.synchronize()
    return s.elapsed_time(e) / n, loss.item()
for _ in range(3): step.step(x, y)
print("eager  ms/step %.2f loss %.5f" % timeit())
t0 = time.time(); step.capture(B, T); print("capture took %.1fs" % (time.time() - t0))
print("graph  ms/step %.2f loss %.5f" % timeit())
print("graph  ms/step %.2f loss %.5f" % timeit())
