# A/B of the planner kernel's launch shape (compile-time PLAN_THREADS x PLAN_MIN_BLOCKS)
for cfg in "128 5" "96 5" "160 3" "128 4"; do
  set -- $cfg
  CRLK_DEFS="PLAN_THREADS=$1 PLAN_MIN_BLOCKS=$2" python -m crl_kino_b200.build --force > /dev/null
  for pq in 1024 8192; do
    timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu --plan-queries $pq > gpurun_out/bench_mb.json 2>gpurun_out/bench_mb.err
    python -c "
import json; d=json.load(open('gpurun_out/bench_mb.json')); print('PLAN T=$1 MINB=$2 planQ=$pq', round(d['value']), d['kernel_ms_per_step']['extend_dwa'], round(d['planning']['queries_per_sec']), d['planning']['ms'])"
  done
done
python -m crl_kino_b200.build --force > /dev/null
