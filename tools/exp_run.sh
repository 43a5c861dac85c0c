mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q -s > gpurun_out/e49_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/e49_pytest.log
grep -E "sigma|passed|failed|rc=|sign, vector" gpurun_out/e49_pytest.log | tail -20
