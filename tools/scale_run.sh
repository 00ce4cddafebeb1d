# Multi-GPU evidence on one box with the driver's command: python -m torch.distributed.run ... bench.py --gpus N
# (cfg4 weak-scaling line + the cfg5 object with the NCCL result gather), and the reference arm on the same box.
#   gpurun --gpus 8 -- 'bash tools/scale_run.sh 8'      gpurun --gpus 2 -- 'bash tools/scale_run.sh 2'
n=${1:-8}
run() { timeout 800 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500 + n)) bench.py --gpus $n "$@"; }
run > gpurun_out/r2_bench_${n}gpu.json 2> gpurun_out/r2_bench_${n}gpu.err
run --impl reference --steps 3 --warmup 1 > gpurun_out/r2_bench_${n}gpu_reference.json 2> gpurun_out/r2_bench_${n}gpu_reference.err
python - "gpurun_out/r2_bench_${n}gpu.json" <<'PY'
import json, sys
try:
    d = json.loads([l for l in open(sys.argv[1]) if l.startswith("{")][0])
    c = d.get("cfg5") or {}
    print(d["n_gpus"], "GPUs: cfg4", round(d["value"]), "e2e", round(d["e2e"]["value"]), "planning", (d.get("planning") or {}).get("queries_per_sec"),
          "| cfg5", c.get("value"), "e2e", (c.get("e2e") or {}).get("value"), "gather ms", c.get("gather_ms"), "capped", (c.get("config") or {}).get("capped_by_memory"))
except Exception as e:
    print(sys.argv[1], "FAILED", e)
PY
