mkdir -p gpurun_out/r2y
cd /root/repo
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/r2y/pytest_gpu.log 2>&1; echo "pytest rc $?"; tail -15 gpurun_out/r2y/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -3
B="python bench.py --no-cpu-baseline --no-gpu-baseline --steps 20 --warmup 3"
run() { name=$1; shift; env "$@" timeout 200 $B > gpurun_out/r2y/bench_$name.json 2> gpurun_out/r2y/bench_$name.err; python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r2y/bench_$name.json").read().strip().splitlines()[-1])
    print("$name", d["ms_per_step"], d["value"], d["clocks"]["sm_mhz"], d.get("gpu_launches"))
except Exception as e:
    print("$name", "failed", e)
PY
}
run A X=1
run B X=1
timeout 300 python bench.py --no-cpu-baseline --no-gpu-baseline --steps 4 --warmup 3 --breakdown gpurun_out/r2y/breakdown.txt > gpurun_out/r2y/bench_bd.json 2> gpurun_out/r2y/bench_bd.err
head -60 gpurun_out/r2y/breakdown.txt
