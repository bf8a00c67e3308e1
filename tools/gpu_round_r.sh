mkdir -p gpurun_out
( timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29543 bench.py --gpus 8 --steps 10 --warmup 3 ) > gpurun_out/r3r_bench_8gpu.json 2> gpurun_out/r3r_bench_8gpu.err; echo "bench8 rc=$?"
( timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29544 tools/tafl_selfplay_bench.py --game 23 --games 8192 --moves 16 --cpu-seconds 0.2 ) > gpurun_out/r3r_sg_selfplay_8gpu.json 2> gpurun_out/r3r_sg_selfplay_8gpu.err; echo "sg8 rc=$?"
tail -2 gpurun_out/r3r_bench_8gpu.err; cut -c1-250 gpurun_out/r3r_bench_8gpu.json; cut -c1-400 gpurun_out/r3r_sg_selfplay_8gpu.json
