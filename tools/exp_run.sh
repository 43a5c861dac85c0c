mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/e45_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/e45_pytest.log
tail -4 gpurun_out/e45_pytest.log
timeout 600 python bench.py > gpurun_out/e45_bench32.json 2> gpurun_out/e45_bench32.err; python -c "
import json; d=json.load(open('gpurun_out/e45_bench32.json')); print(d['value'], d['roofline']['frac'], d['roofline']['kernel'], d['e2e']['value'])"
timeout 900 python tools/op_bench.py > gpurun_out/e45_ops.jsonl 2> gpurun_out/e45_ops.err; cat gpurun_out/e45_ops.jsonl | cut -c1-200
