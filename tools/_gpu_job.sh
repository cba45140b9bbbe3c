mkdir -p gpurun_out
for X in xpA xpB; do
export BRAWL_CUDA_LIB=$PWD/brawl_b200/libbrawl_$X.so
python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "screened or dense_decomposition_conservation" 2>&1 | tail -1
python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_x.json 2> gpurun_out/bench_x.err
tail -c 300 gpurun_out/bench_x.err
python -c "
import json;d=json.load(open('gpurun_out/bench_x.json'));print('chain lib=$X', d['value'],d['e2e']['value'],d['config']['acceptance'],d['config']['energy_per_atom_start_end_Ry'])"
done
