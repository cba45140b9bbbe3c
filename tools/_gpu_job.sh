mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -4
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_r01_reference.json 2> gpurun_out/bench_ref.err
python bench.py > gpurun_out/bench_r01.json 2> gpurun_out/bench.err
tail -c 400 gpurun_out/bench.err
python -c "
import json;d=json.load(open('gpurun_out/bench_r01.json'));print(d['value'],d['e2e']['value'],d['roofline']['frac'],d['config']['acceptance'],d['config']['energy_per_atom_start_end_Ry'], d['cpu_baseline']['value'], d['gpu_launches'])"
python bench.py --workload replicas --no-cpu-baseline > gpurun_out/bench_r01_replicas.json 2> gpurun_out/bench_rep.err
python -c "
import json;d=json.load(open('gpurun_out/bench_r01_replicas.json'));print('replicas', d['value'],d['e2e']['value'],d['roofline']['frac'])"
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --sweeps 16 --no-cpu-baseline > gpurun_out/b_ncu.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:brw_box_metropolis_word -s 20 -c 1 -o gpurun_out/prof_split -f python bench.py --steps 2 --warmup 3 --sweeps 16 --no-cpu-baseline > gpurun_out/b_ncu2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:brw_energy_tile -s 2 -c 1 -o gpurun_out/prof_etile -f python bench.py --steps 2 --warmup 3 --sweeps 16 --no-cpu-baseline > gpurun_out/b_ncu3.log 2>&1
ls -la gpurun_out | head -40
