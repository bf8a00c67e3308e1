mkdir -p gpurun_out
( timeout 1500 python -m pytest tests/test_forest.py tests/test_pybind_mcts.py tests/test_zz_pybind_tafl_pm.py tests/test_stargambit_search.py -m gpu -x -q ) > gpurun_out/r3o_pytest.log 2>&1; echo "pytest rc=$?"
tail -25 gpurun_out/r3o_pytest.log
