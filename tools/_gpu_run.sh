mkdir -p gpurun_out/r2ak
cd /root/repo
O=gpurun_out/r2ak
nvidia-smi -L | head -8 > $O/gpus.txt
# every GPU alone, all at the same time (no communication): what each device of this box does on its own
for d in 0 1 2 3 4 5 6 7; do CUDA_VISIBLE_DEVICES=$d timeout 200 python bench.py --no-cpu-baseline --no-gpu-baseline --steps 20 --warmup 3 > $O/single_$d.json 2> $O/single_$d.err & done; wait
for n in 8 2; do timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2955$n bench.py --gpus $n --steps 20 --warmup 3 --no-cpu-baseline --no-gpu-baseline > $O/bench_n$n.json 2> $O/bench_n$n.err; done
CUDA_VISIBLE_DEVICES=0 timeout 200 python bench.py --no-cpu-baseline --no-gpu-baseline --steps 20 --warmup 3 > $O/bench_n1.json 2> $O/bench_n1.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r2ak/*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split('/')[-1], d.get("n_gpus"), round(d.get("ms_per_step"),3), round(d.get("value")), (d.get("clocks") or {}).get("sm_mhz"))
    except Exception as e:
        print(f, "failed", e)
PY
