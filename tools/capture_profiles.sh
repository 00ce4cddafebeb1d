# Round evidence in one GPU call: tests, the bench line, the ncu launch list of the same command and
# full captures of the dominant kernels (k_extend_dwa of cfg4; k_nearest at cfg5 tree size; the policy
# GEMM and the fused policy-step kernel of the RL extension).  Summaries: tools/ncu_summary.py (here).
set -x
tag=${1:-r2}
timeout 900 python -m pytest tests -q -m gpu 2>&1 | tail -3
timeout 600 python bench.py > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err || exit 1
timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu --no-extra --sustain-s 0 > gpurun_out/${tag}_plain_a.log 2>&1 &&
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -c 400 --csv --log-file gpurun_out/${tag}_launches.csv python bench.py --steps 3 --warmup 3 --no-cpu --no-extra --sustain-s 0 > gpurun_out/${tag}_ncu_a.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:k_extend_dwa -c 1 -o gpurun_out/${tag}_extend -f python bench.py --steps 1 --warmup 3 --no-cpu --no-extra --sustain-s 0 > gpurun_out/${tag}_ncu_b.log 2>&1
timeout 200 python tools/bench_nearest.py > gpurun_out/${tag}_plain_c.log 2>&1 &&
timeout 400 ncu --set full --clock-control none -k regex:k_nearest -s 4 -c 1 -o gpurun_out/${tag}_nearest -f python tools/bench_nearest.py > gpurun_out/${tag}_ncu_c.log 2>&1
timeout 200 python tools/prof_rl.py 8192 > gpurun_out/${tag}_plain_d.log 2>&1 &&
timeout 400 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:"k_mlp_chain|k_rl_advance" -s 20 -c 10 -o gpurun_out/${tag}_rl -f python tools/prof_rl.py 8192 > gpurun_out/${tag}_ncu_d.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/${tag}_rl_launches.csv python tools/prof_rl.py 8192 > gpurun_out/${tag}_ncu_e.log 2>&1
ls -la gpurun_out/${tag}*
