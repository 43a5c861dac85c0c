mkdir -p gpurun_out
for c in 148 296 4096; do
echo -n "ring(new head) acc32 $c: "
FSC_PBS_VARIANT=ring timeout 100 python tools/prof_pbs.py $c 2 2>&1 | grep pbs | tail -1
done
FSC_PBS_VARIANT=ring timeout 400 python -m pytest tests/test_gpu_pbs.py -m gpu -x -q -k "noise or all_messages or variants" 2>&1 | tail -2
