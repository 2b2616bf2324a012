cd /root/repo
python tools/kprof.py lstm96 --iters 3 2>&1 | tail -1
python tools/kprof.py lstm48 --iters 3 2>&1 | tail -1
ROWS=14 python tools/tc_trace.py lstm96 2>&1 | tail -16
ROWS=8 python tools/tc_trace.py lstm48 2>&1 | tail -9
python bench.py --no-cpu-baseline 2>/dev/null | cut -c1-200
timeout 300 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x -k "lstm" 2>&1 | tail -2
