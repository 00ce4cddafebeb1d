# one-launch MLP chain (k_mlp_chain): tile width A/B, slice count, tests
( timeout 300 python -m pytest tests/test_gpu_nets.py -x -q -m gpu 2>&1 | tail -3
for bn in 128 256; do for s in 1 2; do echo "bn=$bn split=$s"; CRLK_MLP_BN=$bn CRLK_RL_SPLIT=$s timeout 120 python tools/prof_rl.py 8192; done; done
for bn in 128 256; do echo "bn=$bn"; CRLK_MLP_BN=$bn timeout 120 python tools/bench_mlp.py; CRLK_MLP_BN=$bn timeout 120 python tools/bench_mlp.py 2048;  CRLK_MLP_BN=$bn timeout 120 python tools/bench_mlp.py 20000; done
timeout 600 python -m pytest tests/test_gpu_batched_rrt.py tests/test_gpu_examples.py tests/test_gpu_gym.py -x -q -m gpu 2>&1 | tail -3 ) > gpurun_out/r2h_chain.txt 2>&1
CRLK_RL_SPLIT=1 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none --profile-from-start off --csv --log-file gpurun_out/r2h_rl_launches_chain_split1.csv python tools/prof_rl.py 8192 > gpurun_out/r2h_ncu.log 2>&1
cat gpurun_out/r2h_chain.txt
