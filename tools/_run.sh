python -m pytest tests/test_gpu_kernels.py tests/test_gpu_pipeline.py tests/test_gpu_edges.py -x -q 2>&1 | tail -3
python tools/hm_time.py
python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-e2e 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('C2', round(d['value']), round(d['ms_per_step']*1e3,1), 'serial', round(d['serial']['ms_per_step']*1e3,1), 'k1', round(d['kernels']['warp_normalize']['ms']*1e3,1), 'fused', round(d['kernels']['heatmap_fused']['ms']*1e3,1))"
