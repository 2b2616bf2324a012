cd /root/repo
python tools/kprof.py lstm96 --iters 3 2>&1 | tail -1
python tools/kprof.py lstm48 --iters 3 2>&1 | tail -1
timeout 300 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x -k "lstm" 2>&1 | tail -2
python bench.py --no-cpu-baseline 2>/dev/null | cut -c1-200
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches_r1l.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_b.log 2>&1; tail -1 gpurun_out/ncu_b.log | cut -c1-100
