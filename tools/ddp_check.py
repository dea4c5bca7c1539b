"""2+ rank check on real GPUs: data-parallel step == single-process step on the concatenated batch.
Run: torchrun --nproc-per-node 2 tools/ddp_check.py"""
import os, sys
import torch, torch.distributed as dist
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "genomics-lm_b200"))
import bench
from codonlm_b200 import TinyGPT
from codonlm_b200.trainer import TrainStep
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local); dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
def build():
    torch.manual_seed(1)
    m = TinyGPT(vocab_size=68, block_size=256, n_layer=2, n_head=4, n_embd=256, dropout=0.0, label_smoothing=0.05,
                use_sdpa=True, termination_aux=True, multi_offset_targets=[2, 4])
    with torch.no_grad():
        m.tok_emb.weight.mul_(0.02); m.pos_emb.weight.mul_(0.02)
    return m.to(dev).train()
kw = dict(lr=1e-3, lr_embedding=1e-3, weight_decay=0.05, offset_weights={2: 0.5, 4: 0.5}, termination_loss_weight=0.1)
B = 4
data = [bench.synthetic_tokens(B, 256, seed=50 + r) for r in range(world)]
# data-parallel: each rank its own micro-batch
dp = TrainStep(build(), **kw, bucket_mb=1)
GRAPH = os.environ.get("CGPT_DDP_GRAPH", "0") == "1"  # the all-reduces captured in the CUDA graph with the step
if GRAPH:
    dp.step(*[t.to(dev) for t in data[rank]])           # first step eagerly: counts the gradient contributions
    dp.capture(B, 256, allow_collectives=True)
    for it in range(2):
        x, y = data[rank]
        dp.step(x.to(dev), y.to(dev))
else:
    for it in range(3):
        x, y = data[rank]
        dp.step(x.to(dev), y.to(dev))
torch.cuda.synchronize()
# reference: one process, gradient accumulation over the same micro-batches (mean of per-micro-batch means)
ref = TrainStep(build(), **kw, process_group=None)
ref.buckets = []; ref.world = 1
for it in range(3):
    ref.zero_grad()
    for r in range(world):
        x, y = data[r]
        ref.forward_backward(x.to(dev), y.to(dev))
    ref.optimizer_step(micro_batches=world)
torch.cuda.synchronize()
worst = 0.0
for g1, g2 in zip(dp.groups, ref.groups):
    worst = max(worst, ((g1.flat - g2.flat).abs().max() / g2.flat.abs().max()).item())
flat = torch.cat([g.flat for g in dp.groups])
gathered = [torch.empty_like(flat) for _ in range(world)]
dist.all_gather(gathered, flat)
same = all(torch.equal(gathered[0], t) for t in gathered)
if rank == 0:
    print(f"ddp_check: graph={GRAPH} world={world} max rel weight diff vs accumulation reference {worst:.3e}; replicas identical: {same}; "
          f"buckets {[len(b.bounds) for b in dp.buckets]}")
    assert worst < 5e-3 and same
import threading
_k = threading.Timer(20.0, os._exit, args=(0,)); _k.daemon = True; _k.start()
dp._graph = None
torch.cuda.synchronize()
dist.destroy_process_group()
_k.cancel()
