python -m pytest tests -m gpu -x -q 2>&1 | tail -5
ncu --set full --clock-control none --import-source on -k regex:word_kernel -s 6 -c 1 -f -o gpurun_out/prof_r01_word python bench.py --steps 1 --warmup 3 --sweeps 16 --no-cpu-baseline > gpurun_out/ncu_word.log 2>&1
tail -3 gpurun_out/ncu_word.log
