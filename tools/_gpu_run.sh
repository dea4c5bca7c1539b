mkdir -p gpurun_out/r2ac
cd /root/repo
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/r2ac/pytest_gpu.log 2>&1; echo "pytest rc $?"; tail -3 gpurun_out/r2ac/pytest_gpu.log
for v in 0 4 6 12; do echo "LN_FWD_VARIANT=$v"; CGPT_LN_REVERSE=0 CGPT_LN_FWD_VARIANT=$v timeout 120 python tools/ln_probe.py 2>&1 | head -1; done
B="python bench.py --no-cpu-baseline --no-gpu-baseline --steps 20 --warmup 3"
run() { name=$1; shift; env "$@" timeout 200 $B > gpurun_out/r2ac/bench_$name.json 2> gpurun_out/r2ac/bench_$name.err; python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r2ac/bench_$name.json").read().strip().splitlines()[-1])
    print("$name", d["ms_per_step"], d["value"], d["clocks"]["sm_mhz"], d.get("gpu_launches"))
except Exception as e:
    print("$name", "failed", e)
PY
}
run base CGPT_LN_REVERSE=0
run lnv6 CGPT_LN_REVERSE=0 CGPT_LN_FWD_VARIANT=6
run base2 CGPT_LN_REVERSE=0
run lnv6b CGPT_LN_REVERSE=0 CGPT_LN_FWD_VARIANT=6
run rev1 CGPT_LN_REVERSE=1
for w in c4_train; do echo "$w auto"; timeout 200 python bench.py --workload $w --no-cpu-baseline --no-gpu-baseline --steps 30 --warmup 3 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['ms_per_step'], d['value'])"; done
