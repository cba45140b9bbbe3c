mkdir -p gpurun_out
for PM in 0 53 90 130 170; do
export BRW_STEPS_A_PERMILLE=$PM
python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_x.json 2> gpurun_out/bench_x.err
tail -c 300 gpurun_out/bench_x.err
python -c "
import json;d=json.load(open('gpurun_out/bench_x.json'));print('chain permille $PM', d['value'],d['e2e']['value'])"
done
