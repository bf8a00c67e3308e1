mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_stargambit.py -m gpu -x -q ) > gpurun_out/r3q_pytest_sg.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r3q_pytest_sg.log
( timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 10 --warmup 3 ) > gpurun_out/r3q_bench_2gpu.json 2> gpurun_out/r3q_bench_2gpu.err; echo "bench2 rc=$?"
( timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 bench.py --impl reference --gpus 2 --steps 3 --warmup 1 ) > gpurun_out/r3q_bench_2gpu_ref.json 2> gpurun_out/r3q_bench_2gpu_ref.err; echo "ref2 rc=$?"
tail -2 gpurun_out/r3q_bench_2gpu.err; cut -c1-250 gpurun_out/r3q_bench_2gpu.json; cut -c1-200 gpurun_out/r3q_bench_2gpu_ref.json
