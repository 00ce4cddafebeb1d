set -e
CRLK_DEFS=CRLK_PHASE_TIMING python -m crl_kino_b200.build --force > /dev/null
for w in 0; do echo "== WIDE=$w"; CRLK_DWA_WIDE=$w python tools/exp_phases.py; done > gpurun_out/phases_wide.txt 2>&1
python -m crl_kino_b200.build --force > /dev/null
cat gpurun_out/phases_wide.txt
