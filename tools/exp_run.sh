mkdir -p gpurun_out
timeout 900 python tools/op_bench.py > gpurun_out/r01b_ops_n1.jsonl 2> gpurun_out/r01b_ops_n1.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/multi_gpu_ops.py > gpurun_out/r01b_ops_256bit_n2.jsonl 2> gpurun_out/r01b_ops_n2.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r01b_bench_n2.json 2> gpurun_out/r01b_bench_n2.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r01b_bench_ref.json 2> gpurun_out/r01b_bench_ref.err
cut -c1-160 gpurun_out/r01b_ops_n1.jsonl; cut -c1-200 gpurun_out/r01b_ops_256bit_n2.jsonl; tail -3 gpurun_out/r01b_ops_n2.err; head -c 300 gpurun_out/r01b_bench_n2.json; echo; head -c 400 gpurun_out/r01b_bench_ref.json
