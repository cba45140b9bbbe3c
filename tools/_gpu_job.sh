mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q -k "walker_kernels_fast" 2>&1 | tail -8
