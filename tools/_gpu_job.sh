mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
for W in 16 256; do
python tools/wl_multi_gpu.py --windows 8 --walkers $W 2> gpurun_out/wl.err | grep workload > gpurun_out/wl_fast3_n1_w$W.json; tail -c 300 gpurun_out/wl.err; cat gpurun_out/wl_fast3_n1_w$W.json
done
