mkdir -p gpurun_out
L=gpurun_out/e29.log
: > $L
for m in 0 -1 -2 -4; do
echo "=== stream v2 mode $m" >> $L
FSC_PBS_DEBUG_CLOCKS=1 FSC_PBS_VARIANT=stream FSC_PBS_STAGGER=$m timeout 100 python tools/prof_pbs.py 4096 1 2>&1 | grep -E "block 0 step  (384|640)|pbs" >> $L
done
cat $L
