mkdir -p gpurun_out
for N in 4 8; do
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$N"
$TR bench.py --gpus $N --steps 10 --warmup 3 2> gpurun_out/bench_n$N.err | grep '^{' > gpurun_out/bench_r01_n$N.json
tail -c 200 gpurun_out/bench_n$N.err; python -c "
import json;d=json.load(open('gpurun_out/bench_r01_n$N.json'));print('N=$N chain', d['value'],d['e2e']['value'],d['n_gpus'])"
done
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29519"
$TR bench.py --gpus 8 --steps 5 --warmup 3 --workload replicas 2> gpurun_out/bench_n8r.err | grep '^{' > gpurun_out/bench_r01_n8_replicas.json
python -c "
import json;d=json.load(open('gpurun_out/bench_r01_n8_replicas.json'));print('N=8 replicas', d['value'],d['e2e']['value'],d['n_gpus'])"
$TR tools/wl_multi_gpu.py --windows 16 --walkers 256 2> gpurun_out/wl_n8.err | grep workload > gpurun_out/wl_n8_w256_win16.json; tail -c 200 gpurun_out/wl_n8.err; cat gpurun_out/wl_n8_w256_win16.json
