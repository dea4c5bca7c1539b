mkdir -p gpurun_out/r2ag
cd /root/repo
LIB=genomics-lm_b200/codonlm_b200/libcgpt_b200.so
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/r2ag/pytest_gpu.log 2>&1; echo "pytest rc $?"; tail -3 gpurun_out/r2ag/pytest_gpu.log
B="python bench.py --no-cpu-baseline --no-gpu-baseline --steps 20 --warmup 3"
run() { name=$1; shift; env "$@" timeout 200 $B > gpurun_out/r2ag/bench_$name.json 2> gpurun_out/r2ag/bench_$name.err; python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r2ag/bench_$name.json").read().strip().splitlines()[-1])
    print("$name", d["ms_per_step"], d["value"], d["clocks"]["sm_mhz"], d["step_flops"]["frac_of_bf16_burst_peak"])
except Exception as e:
    print("$name", "failed", e)
PY
}
for v in new prev new prev; do cp tools/_prev/libcgpt_$v.so $LIB; run $v$RANDOM X=1; done
for v in new prev; do cp tools/_prev/libcgpt_$v.so $LIB; echo "== $v"; timeout 200 python tools/gemm_probe.py qkv_fwd fc1_fwd_gelu fc2_dgrad_mulaux_colsum proj_fwd_res fc1_wgrad 2>&1 | tail -5; timeout 100 python tools/attn_probe.py attn_bwd 2>&1 | tail -1; timeout 100 python tools/attn_probe.py attn_fwd 2>&1 | tail -1; timeout 60 python tools/ln_probe.py 2>&1 | tail -1; done
cp tools/_prev/libcgpt_new.so $LIB
