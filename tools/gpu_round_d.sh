mkdir -p gpurun_out
( timeout 1200 python -m pytest tests/test_pybind_mcts.py tests/test_zz_pybind_tafl_pm.py tests/test_forest.py -m gpu -x -q ) > gpurun_out/r3d_pytest.log 2>&1; echo "pytest rc=$?"
tail -30 gpurun_out/r3d_pytest.log
