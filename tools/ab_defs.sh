# A/B of a compile-time variant against the default build at equal bench settings, in ONE gpurun call
# and without compiling on the GPU box.
#   usage (here):   tools/ab_defs.sh "DWA_THREADS=128 DWA_MIN_BLOCKS=4" [bench flags...]
# builds crl_kino_b200/libcrlk_variant.so with -D<defs> next to the default libcrlk.so, then runs on the
# box: oracle tests + bench with the default library, the same with the variant swapped in.
set -e
defs="$1"; shift || true
flags="${*:---no-cpu --no-extra --steps 40}"
cd "$(dirname "$0")/.."
python -m crl_kino_b200.build
cp crl_kino_b200/libcrlk.so /tmp/libcrlk_default.so
CRLK_DEFS="$defs" python -m crl_kino_b200.build --force
mv crl_kino_b200/libcrlk.so crl_kino_b200/libcrlk_variant.so
python -m crl_kino_b200.build --force
tag=ab_$(echo "$defs" | tr ' =' '__')
gpurun --timeout 400 -- "
timeout 120 python bench.py $flags > gpurun_out/${tag}_default.json 2> gpurun_out/${tag}_default.err
cp crl_kino_b200/libcrlk_variant.so crl_kino_b200/libcrlk.so
timeout 120 python -m pytest tests/test_gpu_extend.py tests/test_gpu_dwa.py tests/test_gpu_batched_rrt.py -x -q -m gpu 2>&1 | tail -2
timeout 120 python bench.py $flags > gpurun_out/${tag}_variant.json 2> gpurun_out/${tag}_variant.err
python - <<'PY'
import json
for w in ('default', 'variant'):
    d = json.load(open('gpurun_out/${tag}_' + w + '.json'))
    print(w, round(d['value']), 'ext/s', d['ms_per_step'], 'ms/step', (d.get('planning') or {}).get('queries_per_sec'))
PY
"
rm -f crl_kino_b200/libcrlk_variant.so
