mkdir -p gpurun_out
( timeout 900 ncu --set full --clock-control none -k regex:k_sp_search -s 3 -c 1 -f -o gpurun_out/r3j_k_sp_search_sg python tools/tafl_selfplay_bench.py --game 23 --games 8192 --moves 6 --warm 2 --cpu-seconds 0.1 ) > gpurun_out/r3j_ncu_sp.log 2>&1; echo "ncu sp rc=$?"
ls -la gpurun_out/*.ncu-rep | tail -3
