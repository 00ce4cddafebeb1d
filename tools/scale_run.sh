# Multi-GPU evidence on one 8-GPU box: cfg4 weak scaling at 8/4/2 ranks and cfg5 (65,536 queries) at 8 ranks.
run() { n=$1; shift; python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500 + n)) bench.py --gpus $n "$@"; }
run 8 --steps 20 --warmup 3 --no-cpu > gpurun_out/r1j_cfg4_8gpu.json 2> gpurun_out/r1j_cfg4_8gpu.err
run 4 --steps 20 --warmup 3 --no-cpu > gpurun_out/r1j_cfg4_4gpu.json 2> gpurun_out/r1j_cfg4_4gpu.err
run 2 --steps 20 --warmup 3 --no-cpu > gpurun_out/r1j_cfg4_2gpu.json 2> gpurun_out/r1j_cfg4_2gpu.err
python bench.py --steps 20 --warmup 3 --no-cpu > gpurun_out/r1j_cfg4_1gpu.json 2> gpurun_out/r1j_cfg4_1gpu.err
run 8 --workload cfg5 --steps 4 --warmup 1 > gpurun_out/r1j_cfg5_8gpu.json 2> gpurun_out/r1j_cfg5_8gpu.err
for f in gpurun_out/r1j_cfg4_?gpu.json gpurun_out/r1j_cfg5_8gpu.json; do python - "$f" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1], d["n_gpus"], round(d["value"]), round(d["e2e"]["value"]), d.get("planning", {}).get("queries_per_sec"))
except Exception as e:
    print(sys.argv[1], "FAILED", e)
PY
done
tail -3 gpurun_out/r1j_cfg5_8gpu.err
