python -m pytest tests -m gpu -x -q -k "dense or trajectory or planner" 2>&1 | tail -3
python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r01_dense4.json 2> gpurun_out/bench_dense.err
tail -c 400 gpurun_out/bench_dense.err
python -c "
import json;d=json.load(open('gpurun_out/bench_r01_dense4.json'));print(d['value'],d['e2e']['value'],d['roofline']['frac'],d['config']['acceptance'],d['config']['energy_per_atom_start_end_Ry'])"
