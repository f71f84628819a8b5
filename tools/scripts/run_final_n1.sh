# final single-GPU evidence of a round: full GPU suite, smoke, default bench, a >= 2 s timed region, the reference arm
set -x
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/final_pytest_gpu.log 2>&1; tail -3 gpurun_out/final_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('SMOKE_OK')" > gpurun_out/final_smoke.log 2>&1; tail -2 gpurun_out/final_smoke.log
timeout 900 python bench.py > gpurun_out/final_bench_default.json 2> gpurun_out/final_bench_default.err
timeout 600 python bench.py --steps 400 --no-cpu-baseline --no-extra-configs --profile-ops > gpurun_out/final_bench_c2_400steps.json 2> gpurun_out/final_per_op_c2.txt
timeout 900 python bench.py --impl reference > gpurun_out/final_bench_reference_arm.json 2> gpurun_out/final_bench_reference_arm.err
python - <<PY
import json
for f in ("final_bench_default", "final_bench_c2_400steps", "final_bench_reference_arm"):
    d = json.loads(open("gpurun_out/%s.json" % f).read().strip().splitlines()[-1])
    print(f, d.get("value"), d.get("ms_per_step"), (d.get("e2e") or {}).get("value"), (d.get("clocks") or {}).get("sm_mhz"),
          (d.get("roofline") or {}).get("frac"), (d.get("cpu_baseline") or {}).get("value"))
    for k, v in (d.get("configs") or {}).items():
        print("   ", k, v.get("value"), v.get("ms_per_step"))
PY
