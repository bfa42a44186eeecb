nvidia-smi topo -m > gpurun_out/r2_topo.txt 2>&1
python -m pytest tests/test_gpu_multi.py -x -q 2>&1 | tail -4
for n in 4 8; do python bench.py --gpus $n --steps 20 --warmup 3 > gpurun_out/r2_bench_n$n.json 2> gpurun_out/r2_bench_n$n.err; python -c "
import json; d=json.loads(open('gpurun_out/r2_bench_n$n.json').read().strip().splitlines()[-1]); print($n, 'value', round(d['value']), 'ms/step', d['ms_per_step'], 'serial', round(d['serial']['value']), 'allreduce_us', d['allreduce_us'], 'per_rank', [round(x,3) for x in d['per_rank_ms_per_region']]); e=d['e2e']; print('  e2e', round(e['value']), 'GB/s/gpu', round(e['h2d_GBps_per_gpu'],1), 'ceiling', [round(x,1) for x in e['h2d_ceiling_GBps_per_gpu']], 'frac', round(e['frac_of_h2d_ceiling'],3)); print('  check', d['check'])"; done
for n in 4 8; do python bench.py --gpus $n --workload C4 --steps 3 > gpurun_out/r2_c4_n$n.json 2> gpurun_out/r2_c4_n$n.err; python -c "
import json; d=json.loads(open('gpurun_out/r2_c4_n$n.json').read().strip().splitlines()[-1]); print('C4', $n, round(d['value']), d['ms_per_step'], d['check']['n_wrong'], d['check']['curve_hist_sha1'], d['check']['curve_n_valid'], d['check']['epe_sum'])"; done
