# usage: bash tools/_gpu_jobN.sh N  -- the default bench under torchrun on N GPUs of one box
N=$1
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$N"
timeout 500 $TR bench.py --gpus $N --steps 10 --warmup 3 2> gpurun_out/bench_r02_n$N.err | grep '^{' > gpurun_out/bench_r02_n$N.json
tail -c 300 gpurun_out/bench_r02_n$N.err
python -c "
import json;d=json.load(open('gpurun_out/bench_r02_n$N.json'));print('N=$N chain', d['value'],d['e2e']['value'],d['e2e']['with_reference_config_grid']['value'],d['n_gpus']); w=d['extra']['wl']; print('wl', w['value'], w['windows'], w['nrmse_vs_reference_golden'], sum(w['sweeps_calls_per_stage'])); print('ns', d['extra']['ns']['value'])"
