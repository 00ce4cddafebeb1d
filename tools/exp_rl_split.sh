# RL extension pass (8192 queries): slice count A/B and warm-cache per-kernel durations
for s in 1 2 4; do echo "split=$s"; CRLK_RL_SPLIT=$s timeout 120 python tools/prof_rl.py 8192; done > gpurun_out/r2d_rl_split.txt 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none --profile-from-start off --csv --log-file gpurun_out/r2d_rl_launches_warm.csv python tools/prof_rl.py 8192 > gpurun_out/r2d_ncu.log 2>&1
CRLK_RL_SPLIT=1 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none --profile-from-start off --csv --log-file gpurun_out/r2d_rl_launches_warm_split1.csv python tools/prof_rl.py 8192 > gpurun_out/r2d_ncu1.log 2>&1
cat gpurun_out/r2d_rl_split.txt
