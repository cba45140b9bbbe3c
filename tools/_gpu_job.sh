mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "dense or screened or planner or word or production or full_size" 2>&1 | tail -3
for L in 0 2; do
python bench.py --steps 10 --warmup 3 --no-cpu-baseline --layout $L > gpurun_out/bench_x.json 2> gpurun_out/bench_x.err
tail -c 300 gpurun_out/bench_x.err
python -c "
import json;d=json.load(open('gpurun_out/bench_x.json'));print('chain layout $L', d['value'],d['e2e']['value'],d['config']['acceptance'],d['config']['energy_per_atom_start_end_Ry'])"
done
python bench.py --steps 5 --warmup 3 --no-cpu-baseline --workload replicas > gpurun_out/bench_xr.json 2> gpurun_out/bench_xr.err
python -c "
import json;d=json.load(open('gpurun_out/bench_xr.json'));print('replicas', d['value'],d['e2e']['value'])"
