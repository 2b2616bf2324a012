cd /root/repo
for shape in enc0_ftb2 enc0_conv enc0_dc_c2 enc0_rw; do
  for cfg in "0 0" "2 2" "3 2" "4 2" "6 1" "8 1" "4 1"; do
    set -- $cfg
    if [ "$1" = "0" ]; then echo -n "default      "; python tools/kprof.py $shape --iters 4 2>&1 | tail -1;
    else echo -n "stages $1 per_sm $2  "; AERO_TC_STAGES=$1 AERO_TC_PER_SM=$2 python tools/kprof.py $shape --iters 4 2>&1 | tail -1; fi
  done
  echo -n "ungrouped    "; AERO_TC_GROUPED_BN=0 python tools/kprof.py $shape --iters 4 2>&1 | tail -1
done
