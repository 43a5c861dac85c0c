mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 tools/sign_bench.py > gpurun_out/r01b_sign_n8.jsonl 2> gpurun_out/r01b_sign_n8.err; echo rc=$?
grep summary gpurun_out/r01b_sign_n8.jsonl | cut -c1-330; grep -c matches_reference gpurun_out/r01b_sign_n8.jsonl
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus 8 --steps 5 --warmup 3 > gpurun_out/r01b_bench_n8.json 2> gpurun_out/r01b_bench_n8.err; echo rc=$?
head -c 330 gpurun_out/r01b_bench_n8.json; echo
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29523 tools/multi_gpu_ops.py > gpurun_out/r01b_ops_256bit_n8.jsonl 2> gpurun_out/r01b_ops_n8.err; echo rc=$?
cut -c1-200 gpurun_out/r01b_ops_256bit_n8.jsonl
