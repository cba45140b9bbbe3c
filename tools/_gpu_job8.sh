mkdir -p gpurun_out
N=8
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29518"
timeout 500 $TR bench.py --gpus $N --steps 10 --warmup 3 2> gpurun_out/bench_r02_n$N.err | grep '^{' > gpurun_out/bench_r02_n$N.json
tail -c 300 gpurun_out/bench_r02_n$N.err
python -c "
import json;d=json.load(open('gpurun_out/bench_r02_n$N.json'));print('N=$N chain', d['value'],d['e2e']['value'],d['n_gpus']); w=d['extra']['wl']; print('wl', w['value'], w['windows'], w['nrmse_vs_reference_golden'], w['sweeps_calls_per_stage']); print(w['host_seconds_per_rank'][0])"
