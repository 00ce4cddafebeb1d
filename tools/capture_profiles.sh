# Round evidence: tests, bench, ncu launch list and full captures of the two dominant kernels.
set -x
tag=${1:-r1h}
timeout 1100 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
timeout 400 python bench.py > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err || exit 1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -c 400 --csv --log-file gpurun_out/${tag}_launches.csv python bench.py --steps 3 --warmup 3 --no-cpu > gpurun_out/${tag}_ncu_a.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:k_extend_dwa -c 1 -o gpurun_out/${tag}_extend -f python bench.py --steps 1 --warmup 3 --no-cpu --no-extra > gpurun_out/${tag}_ncu_b.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:k_plan_dwa -c 1 -o gpurun_out/${tag}_plan -f python bench.py --steps 1 --warmup 3 --no-cpu > gpurun_out/${tag}_ncu_c.log 2>&1
ls -la gpurun_out/${tag}*
