mkdir -p gpurun_out
timeout 1200 compute-sanitizer --tool racecheck --racecheck-report all --print-limit 400 python tools/racecheck_box.py > gpurun_out/racecheck.log 2>&1; echo "racecheck rc=$?"; tail -12 gpurun_out/racecheck.log
python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "dense or screened or planner or word or production" 2>&1 | tail -2
python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_x.json 2> gpurun_out/bench_x.err
python -c "
import json;d=json.load(open('gpurun_out/bench_x.json'));print('chain', d['value'],d['e2e']['value'],d['config']['acceptance'],d['config']['energy_per_atom_start_end_Ry'])"
