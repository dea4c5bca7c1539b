mkdir -p gpurun_out/r2aa
cd /root/repo
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/r2aa/pytest_gpu.log 2>&1; echo "pytest rc $?"; tail -5 gpurun_out/r2aa/pytest_gpu.log
B="python bench.py --no-cpu-baseline --no-gpu-baseline --steps 20 --warmup 3"
run() { name=$1; shift; env "$@" timeout 200 $B > gpurun_out/r2aa/bench_$name.json 2> gpurun_out/r2aa/bench_$name.err; python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r2aa/bench_$name.json").read().strip().splitlines()[-1])
    print("$name", d["ms_per_step"], d["value"], d["clocks"]["sm_mhz"], d.get("gpu_launches"))
except Exception as e:
    print("$name", "failed", e)
PY
tail -2 gpurun_out/r2aa/bench_$name.err
}
run pdl1 X=1
run pdl0 CGPT_PDL=0
run pdl1b X=1
run pdl0b CGPT_PDL=0
