mkdir -p gpurun_out
( timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 2 --steps 5 --warmup 3 --no-cpu-baseline ) > gpurun_out/r15_bench_2gpu.json 2> gpurun_out/r15_bench_2gpu.err; echo "rc=$?" >> gpurun_out/r15_bench_2gpu.err
python -c "
import json
for l in open('gpurun_out/r15_bench_2gpu.json'):
    if l.startswith('{'):
        d=json.loads(l); print(d['n_gpus'], d['value'], d['e2e']['value'], d['e2e_nn_device']['value'], d['clocks'])"; tail -n 4 gpurun_out/r15_bench_2gpu.err
