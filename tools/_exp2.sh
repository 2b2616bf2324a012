cd /root/repo
timeout 600 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_parity.py -m gpu -q -x 2>&1 | tail -3
python bench.py --no-cpu-baseline 2>/dev/null | cut -c1-200
for v in "1 0" "0 0" "1 1" "0 1"; do set -- $v; echo "direct_f16=$1 direct_f32=$2"; AERO_TC_DIRECT_F16=$1 AERO_TC_DIRECT_F32=$2 python bench.py --no-cpu-baseline 2>/dev/null | cut -c80-200; done
for s in enc0_ftb2 enc0_conv enc0_rw enc1_ftb2 dec3_rw; do for d in 1 0; do echo -n "f16 direct=$d "; AERO_TC_DIRECT_F16=$d python tools/kprof.py $s --iters 4 2>&1 | tail -1; done; done
for s in enc0_dc_c2 dec0_rw dec1_rw; do for d in 1 0; do echo -n "f32 direct=$d "; AERO_TC_DIRECT_F32=$d python tools/kprof.py $s --iters 4 2>&1 | tail -1; done; done
