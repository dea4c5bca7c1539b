mkdir -p gpurun_out/r2x
cd /root/repo
timeout 300 python -m pytest tests/test_kernels_gpu.py -m gpu -x -q -k "layernorm" > gpurun_out/r2x/pytest_ln.log 2>&1; echo "pytest rc $?"
B="python bench.py --no-cpu-baseline --no-gpu-baseline --steps 20 --warmup 3"
run() { name=$1; shift; env "$@" timeout 200 $B > gpurun_out/r2x/bench_$name.json 2> gpurun_out/r2x/bench_$name.err; python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r2x/bench_$name.json").read().strip().splitlines()[-1])
    print("$name", d["ms_per_step"], d["value"], d["clocks"]["sm_mhz"])
except Exception as e:
    print("$name", "failed", e)
PY
}
run A_base CGPT_LN_STREAM=0 CGPT_LN_REVERSE=0 CGPT_GEMM_DEBUG=8
run B_gemm CGPT_LN_STREAM=0 CGPT_LN_REVERSE=0
run C_stream CGPT_LN_REVERSE=0
run D_all CGPT_LN_STREAM=1
run E_fwdstream CGPT_LN_STREAM=2
run F_rev_only CGPT_LN_STREAM=0
run A2_base CGPT_LN_STREAM=0 CGPT_LN_REVERSE=0 CGPT_GEMM_DEBUG=8
run D2_all CGPT_LN_STREAM=1
