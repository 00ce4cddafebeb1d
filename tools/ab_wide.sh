# A/B of the two DWA launch shapes: parity tests, bench, per-phase cycles.
for w in 0; do
  echo "== WIDE=$w"
  CRLK_DWA_WIDE=$w timeout 600 python -m pytest tests/test_gpu_dwa.py tests/test_gpu_extend.py tests/test_gpu_batched_rrt.py -x -q -m gpu 2>&1 | tail -3
  CRLK_DWA_WIDE=$w timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu > gpurun_out/bench_wide$w.json 2>gpurun_out/bench_wide$w.err
  python -c "
import json; d=json.load(open('gpurun_out/bench_wide$w.json')); print(d['value'], d['e2e']['value'], d['kernel_ms_per_step'], d['planning']['queries_per_sec'], d['planning']['ms'])"
done
bash tools/run_phases.sh
