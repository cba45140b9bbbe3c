for v in BASE ROWTAB F2I32 BOTH; do
  cp brawl_b200/libbrawl_cuda_$v.so brawl_b200/libbrawl_cuda.so
  python bench.py --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/ab_$v.json 2> gpurun_out/ab.err
  python -c "
import json;d=json.load(open('gpurun_out/ab_$v.json'));print('$v', d['value'],d['e2e']['value'],d['config']['acceptance'],d['config']['energy_per_atom_start_end_Ry'][1])"
done
