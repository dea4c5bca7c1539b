mkdir -p gpurun_out/r2af
cd /root/repo
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/r2af/pytest_gpu.log 2>&1; echo "pytest rc $?"; tail -3 gpurun_out/r2af/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
B="python bench.py --no-cpu-baseline --no-gpu-baseline --steps 20 --warmup 3"
run() { name=$1; shift; env "$@" timeout 200 $B > gpurun_out/r2af/bench_$name.json 2> gpurun_out/r2af/bench_$name.err; python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r2af/bench_$name.json").read().strip().splitlines()[-1])
    print("$name", d["ms_per_step"], d["value"], d["clocks"]["sm_mhz"], d.get("gpu_launches"), d["step_flops"]["frac_of_bf16_burst_peak"])
except Exception as e:
    print("$name", "failed", e)
PY
}
run a X=1
run b X=1
