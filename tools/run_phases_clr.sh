set -e
CRLK_DEFS=CRLK_PHASE_TIMING python -m crl_kino_b200.build --force > /dev/null
USE_CLEARANCE=1 python tools/exp_phases.py > gpurun_out/phases_clr.txt 2>&1 || true
python -m crl_kino_b200.build --force > /dev/null
cat gpurun_out/phases_clr.txt
