mkdir -p gpurun_out
L=gpurun_out/e47.log
: > $L
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/e47_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/e47_pytest.log
tail -3 gpurun_out/e47_pytest.log >> $L
for v in ring stream; do for c in 148 296 4096; do
echo -n "$v acc32 $c: " >> $L
FSC_PBS_VARIANT=$v timeout 100 python tools/prof_pbs.py $c 2 2>&1 | grep pbs | tail -1 >> $L
done; done
for c in 148 296 4096; do
echo -n "ring acc64 $c: " >> $L
FSC_BENCH_ACC_BITS=64 timeout 100 python tools/prof_pbs.py $c 2 2>&1 | grep pbs | tail -1 >> $L
done
cat $L
