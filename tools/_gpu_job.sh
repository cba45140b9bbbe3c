mkdir -p gpurun_out
for X in "" xpB xpC xpD; do
if [ -n "$X" ]; then export BRAWL_CUDA_LIB=$PWD/brawl_b200/libbrawl_$X.so; else unset BRAWL_CUDA_LIB; fi
python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "dense or screened or planner or word" 2>&1 | tail -1
for L in 0 2; do
python bench.py --steps 5 --warmup 3 --no-cpu-baseline --layout $L > gpurun_out/bench_L$L.json 2> gpurun_out/bench_L$L.err
tail -c 300 gpurun_out/bench_L$L.err
python -c "
import json;d=json.load(open('gpurun_out/bench_L$L.json'));print('chain lib=$X layout $L', d['value'],d['e2e']['value'],d['config']['acceptance'],d['config']['energy_per_atom_start_end_Ry'])"
done
done
unset BRAWL_CUDA_LIB
for SP in 236 314; do
python bench.py --steps 5 --warmup 3 --no-cpu-baseline --steps-per-phase $SP > gpurun_out/bench_sp.json 2> gpurun_out/bench_sp.err
python -c "
import json;d=json.load(open('gpurun_out/bench_sp.json'));print('chain steps/phase $SP', d['value'],d['e2e']['value'],d['config']['acceptance'],d['config']['energy_per_atom_start_end_Ry'])"
done
