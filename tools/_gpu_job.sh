mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_x.json 2> gpurun_out/bench_x.err
tail -c 300 gpurun_out/bench_x.err
python -c "
import json;d=json.load(open('gpurun_out/bench_x.json'));print('chain', d['value'],d['e2e']['value'],d['config']['acceptance'],d['config']['energy_per_atom_start_end_Ry'])"
