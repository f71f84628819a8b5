# weak scaling at 8 GPUs with / without the per-step all-gather of the statistics (same box, back to back)
for v in 0 1 0 1; do
  BNN_BENCH_WEAK_GATHER=$v timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 2966$v bench.py --gpus 8 --steps 40 --warmup 5 --no-extra-configs --no-cpu-baseline > gpurun_out/r03h_n8_gather$v.json 2> gpurun_out/r03h_n8_gather$v.err
  python - <<PY
import json
d=json.loads(open("gpurun_out/r03h_n8_gather$v.json").read().strip().splitlines()[-1])
print("gather=$v", round(d["value"]), d["ms_per_step"], [round(x,3) for x in d["per_rank_ms_per_step"]])
PY
done
timeout 300 python bench.py --steps 40 --warmup 5 --no-extra-configs --no-cpu-baseline > gpurun_out/r03h_n1.json 2>/dev/null
python - <<PY
import json
d=json.loads(open("gpurun_out/r03h_n1.json").read().strip().splitlines()[-1])
print("n1", round(d["value"]), d["ms_per_step"])
PY
