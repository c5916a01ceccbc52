timeout 120 python tools/factor_check.py 0 lu 2>&1 | tail -3
timeout 100 python tools/lu_panel.py 2>&1 | head -3
timeout 100 python tools/factor_timing.py lu
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "lu" 2>&1 | tail -2
