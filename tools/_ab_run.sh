timeout 900 python -m pytest tests -m gpu -x -q --durations=3 > gpurun_out/r02o_pytest_gpu.log 2>&1
tail -8 gpurun_out/r02o_pytest_gpu.log
python bench.py > gpurun_out/r02o_bench_n1.json 2> gpurun_out/r02o_bench_n1.err
python - <<'P'
import json
d=json.loads(open('gpurun_out/r02o_bench_n1.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['sync']['value'], d['roofline']['frac'], d['clocks'])
for k,c in d['configs'].items(): print(k, c['ms_per_step'], c['value'], c.get('frame_matches_reference'), c['roofline']['frac'])
P
