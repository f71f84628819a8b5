rm -f gpurun_out/multi_gpu_report.jsonl
timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -x -q > gpurun_out/r02_multi_gpu_8ranks_final.log 2>&1
tail -3 gpurun_out/r02_multi_gpu_8ranks_final.log
cp gpurun_out/multi_gpu_report.jsonl gpurun_out/r02_multi_gpu_report_8ranks_final.jsonl
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29621 bench.py --gpus 8 --steps 30 --warmup 3 > gpurun_out/r02_bench_n8_final.json 2> gpurun_out/r02_bench_n8_final.err
python - <<PY
import json
d=json.loads(open("gpurun_out/r02_bench_n8_final.json").read().strip().splitlines()[-1])
print(d["value"], d["ms_per_step"], d["per_rank_ms_per_step"])
c=d["c5_strong"]; print({k:c[k] for k in ("value","ms_per_step","speedup_vs_n1","efficiency_vs_n1","collective_us","collective_us_nccl","collective","n1_ms_per_step")})
PY
