python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
python bench.py > gpurun_out/bench_r01_final2.json 2> gpurun_out/bench_final.err; tail -c 300 gpurun_out/bench_final.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_r01_reference2.json 2>> gpurun_out/bench_final.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r01d.csv python bench.py --steps 2 --warmup 3 --sweeps 16 --no-cpu-baseline > gpurun_out/ncu_launch.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:word_kernel -s 6 -c 1 -f -o gpurun_out/prof_r01_dense28 python bench.py --steps 1 --warmup 3 --sweeps 16 --no-cpu-baseline > gpurun_out/ncu_dense28.log 2>&1
tail -2 gpurun_out/ncu_dense28.log
python -c "
import json
for f in ('gpurun_out/bench_r01_final2.json','gpurun_out/bench_r01_reference2.json'):
    d=json.load(open(f)); print(f, d['value'], d.get('e2e'), d.get('cpu_baseline'), d.get('clocks'))"
