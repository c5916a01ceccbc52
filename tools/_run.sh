timeout 120 python tools/factor_check.py 0 lu 2>&1 | tail -6
echo "rc=$?"
timeout 100 python tools/lu_panel.py 2>&1 | tail -8
timeout 100 python tools/factor_timing.py lu
