mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -5
python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_x.json 2> gpurun_out/bench_x.err
tail -c 300 gpurun_out/bench_x.err
python -c "
import json;d=json.load(open('gpurun_out/bench_x.json'));print('chain', d['value'],d['e2e']['value'],d['config']['acceptance'],d['config']['energy_per_atom_start_end_Ry'])"
python bench.py --steps 5 --warmup 3 --no-cpu-baseline --workload replicas > gpurun_out/bench_xr.json 2> gpurun_out/bench_xr.err
python -c "
import json;d=json.load(open('gpurun_out/bench_xr.json'));print('replicas', d['value'],d['e2e']['value'])"
ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/launches_x.csv python bench.py --steps 2 --warmup 3 --sweeps 16 --no-cpu-baseline > gpurun_out/b_ncu.log 2>&1
grep -E "energy|pack|tree" gpurun_out/launches_x.csv | awk -F'","' '{print substr($5,1,40), $NF}' | sort | uniq -c | sort -k2 | head -12
