mkdir -p gpurun_out
L=gpurun_out/e43.log
: > $L
for o in "2000,1200" "0,0" "4500,2000" "2000,400"; do
echo -n "offsets $o: " >> $L
FSC_PBS_VARIANT=stream FSC_PBS_OFFSETS=$o timeout 100 python tools/prof_pbs.py 4096 1 2>&1 | grep -E "pbs" >> $L
done
for c in 100 148 296; do
FSC_PBS_VARIANT=stream timeout 100 python tools/prof_pbs.py $c 2 2>&1 | grep pbs | tail -1 >> $L
done
cat $L
