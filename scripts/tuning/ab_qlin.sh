cd /root/repo
run() { echo -n "$1 "; shift; env "$@" python scripts/profile_point.py --H 15 --C 2 --B 1048576 --reps 4; }
run "H15 stock(lin q104)" OCD_KERNEL_FORM=wide
for v in qnolin qlin112 qlin96; do run "H15 $v" OCD_B200_LIB=scratch/libocd_$v.so OCD_KERNEL_FORM=wide; done
python scripts/profile_point.py --H 50 --C 2 --B 1048576 --reps 3
python -m pytest tests -m gpu -q -x 2>&1 | tail -3
