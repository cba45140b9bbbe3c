python -m pytest tests -m gpu -x -q 2>&1 | tail -6
python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r01_pair.json 2> gpurun_out/bench_pair.err
tail -c 400 gpurun_out/bench_pair.err
python -c "
import json;d=json.load(open('gpurun_out/bench_r01_pair.json'));print(d['value'],d['e2e']['value'],d['roofline']['frac'],d['config']['acceptance'],d['config']['energy_per_atom_start_end_Ry'])"
