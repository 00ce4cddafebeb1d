set -e
CRLK_DEFS=CRLK_PHASE_TIMING python -m crl_kino_b200.build --force > /dev/null
( echo "== standard kernel (WIDE=0)"; CRLK_DWA_WIDE=0 PHASE_Q=1,148,1776 python tools/exp_phases.py ) > gpurun_out/phases_pair.txt 2>&1
python -m crl_kino_b200.build --force > /dev/null
cat gpurun_out/phases_pair.txt
