# RL extension pass (8192 queries): GEMM kernel choice (single-CTA persistent vs CTA-pair tiles) x slice count
timeout 300 python -m pytest tests/test_gpu_nets.py -x -q -m gpu 2>&1 | tail -3 > gpurun_out/r2f_rl_pair.txt
CRLK_MLP_PAIR=256 timeout 300 python -m pytest tests/test_gpu_nets.py -x -q -m gpu 2>&1 | tail -3 >> gpurun_out/r2f_rl_pair.txt
for p in 0 256; do for s in 1 2; do echo "pair=$p split=$s"; CRLK_MLP_PAIR=$p CRLK_RL_SPLIT=$s timeout 120 python tools/prof_rl.py 8192; done; done >> gpurun_out/r2f_rl_pair.txt 2>&1
timeout 120 python tools/bench_mlp.py >> gpurun_out/r2f_rl_pair.txt 2>&1
CRLK_RL_SPLIT=1 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none --profile-from-start off --csv --log-file gpurun_out/r2f_rl_launches_split1.csv python tools/prof_rl.py 8192 > gpurun_out/r2f_ncu.log 2>&1
CRLK_MLP_PAIR=256 CRLK_RL_SPLIT=1 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none --profile-from-start off --csv --log-file gpurun_out/r2f_rl_launches_pair256_split1.csv python tools/prof_rl.py 8192 > gpurun_out/r2f_ncu2.log 2>&1
cat gpurun_out/r2f_rl_pair.txt
