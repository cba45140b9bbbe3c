mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
for W in 64 256; do
python tools/wl_multi_gpu.py --windows 8 --walkers $W > gpurun_out/wl_n1_w$W.json 2> gpurun_out/wl_n1.err; tail -c 300 gpurun_out/wl_n1.err; cat gpurun_out/wl_n1_w$W.json
done
$TR tools/wl_multi_gpu.py --windows 8 --walkers 256 2> gpurun_out/wl_n2.err | grep workload > gpurun_out/wl_n2_w256.json; tail -c 300 gpurun_out/wl_n2.err; cat gpurun_out/wl_n2_w256.json
$TR tools/wl_multi_gpu.py --windows 16 --walkers 256 2> gpurun_out/wl_n2.err | grep workload > gpurun_out/wl_n2_w256_win16.json; tail -c 300 gpurun_out/wl_n2.err; cat gpurun_out/wl_n2_w256_win16.json
