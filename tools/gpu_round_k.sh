mkdir -p gpurun_out
rm -f gpurun_out/r3k_lock.jsonl
for lock in 0 1; do
for spec in "23 8192" "20 8192" "0 8192" "0 16384" "1 4096"; do
set -- $spec
( B2AZ_SP_LOCKSTEP=$lock timeout 600 python tools/tafl_selfplay_bench.py --game $1 --games $2 --moves 16 --cpu-seconds 0.5 ) 2>> gpurun_out/r3k_lock.err | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print(json.dumps({'lockstep': $lock, 'game': $1, 'games': $2, 'sims_per_s': d['simulations_per_second'], 'device_sims_per_s': d['simulations_per_second_device'], 'moves_per_s': d['moves_per_second']}))
" >> gpurun_out/r3k_lock.jsonl
done
done
cat gpurun_out/r3k_lock.jsonl; tail -3 gpurun_out/r3k_lock.err
( B2AZ_SP_LOCKSTEP=1 timeout 900 python -m pytest tests/test_stargambit_search.py tests/test_tafl_selfplay.py -m gpu -x -q ) > gpurun_out/r3k_pytest_lock.log 2>&1; echo "pytest(lock=1) rc=$?"; tail -3 gpurun_out/r3k_pytest_lock.log
