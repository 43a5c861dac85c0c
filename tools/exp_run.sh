mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/e48_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/e48_pytest.log
tail -3 gpurun_out/e48_pytest.log
for v in ring stream; do for c in 148 4096; do
echo -n "$v acc32 $c: "
FSC_PBS_VARIANT=$v timeout 100 python tools/prof_pbs.py $c 2 2>&1 | grep pbs | tail -1
done; done
echo -n "ring acc64 4096: "; FSC_BENCH_ACC_BITS=64 timeout 100 python tools/prof_pbs.py 4096 2 2>&1 | grep pbs | tail -1
python -c "import __graft_entry__ as g; g.smoke()"
