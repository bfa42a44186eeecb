python -m pytest tests/test_gpu_kernels.py tests/test_gpu_pipeline.py tests/test_gpu_edges.py -x -q 2>&1 | tail -3
python tools/hm_time.py
for w in C2 C3 C5; do python bench.py --workload $w --steps 50 --no-cpu-baseline --no-e2e 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$w', round(d['value']), d['ms_per_step'], d['serial']['ms_per_step'], d['kernels'])"; done
