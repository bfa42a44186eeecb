python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python bench.py --steps 20 --warmup 3 > gpurun_out/r2_final_bench.json 2> gpurun_out/r2_final_bench.err; python -c "
import json; d=json.loads(open('gpurun_out/r2_final_bench.json').read().strip().splitlines()[-1]); print('C2 value', round(d['value']), d['ms_per_step'], 'serial', round(d['serial']['value']), 'frac', d['roofline']['frac'], 'alone', d['kernels']['alone']); print(' e2e', round(d['e2e']['value']), d['e2e']['frac_of_h2d_ceiling'], d['e2e']['jpeg'].get('value'), 'cpu', d['cpu_baseline']['value'], d['cpu_baseline']['kind'], d['clocks'])"
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
