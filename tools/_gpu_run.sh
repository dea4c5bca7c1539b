mkdir -p gpurun_out/r2aj
cd /root/repo
O=gpurun_out/r2aj
timeout 600 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "pytest rc $?"; tail -3 $O/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
( time timeout 600 python bench.py > $O/bench_c3.json 2> $O/bench_c3.err ) 2>&1 | grep real
timeout 300 python bench.py --impl reference --steps 4 --warmup 1 > $O/bench_ref.json 2> $O/bench_ref.err; tail -c 300 $O/bench_ref.json
for w in c2 c4_train c4_infer c5_attn; do timeout 300 python bench.py --workload $w --no-cpu-baseline --no-gpu-baseline --steps 30 --warmup 3 > $O/bench_$w.json 2> $O/bench_$w.err; done
timeout 300 python bench.py --dropout 0.1 --no-cpu-baseline --no-gpu-baseline --steps 12 --warmup 3 > $O/bench_c3_dropout.json 2> $O/bench_c3_dropout.err
timeout 300 python bench.py --tokens realistic --no-cpu-baseline --no-gpu-baseline --steps 12 --warmup 3 > $O/bench_c3_realistic.json 2> $O/bench_c3_realistic.err
timeout 300 python bench.py --no-graph --no-cpu-baseline --no-gpu-baseline --steps 8 --warmup 3 > $O/bench_c3_eager.json 2> $O/bench_c3_eager.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r2aj/bench_*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split('/')[-1], d.get("ms_per_step"), d.get("value"), (d.get("clocks") or {}).get("sm_mhz"), (d.get("step_flops") or {}).get("frac_of_bf16_burst_peak"), (d.get("e2e") or {}).get("value"))
    except Exception as e:
        print(f, "failed", e)
PY
CUDA_VISIBLE_DEVICES=0 timeout 400 python tools/train_cli_check.py /tmp/cli_check > $O/train_cli.log 2>&1; echo "cli rc $?"; tail -4 $O/train_cli.log
