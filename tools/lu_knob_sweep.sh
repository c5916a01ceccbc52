python tools/lu_sweep.py 16384 2
NAB_LU_LEAF=32 python tools/lu_sweep.py 16384 2
NAB_LU_NB=384 python tools/lu_sweep.py 16384 2
NAB_LU_NB=640 python tools/lu_sweep.py 16384 2
NAB_LU_NB=768 python tools/lu_sweep.py 16384 2
NAB_LU_RPCAP=64 python tools/lu_sweep.py 16384 2
NAB_LU_RPCAP=132 python tools/lu_sweep.py 16384 2
NAB_LU_TP0=1.6e-6 NAB_LU_TP1=1.0e-6 python tools/lu_sweep.py 16384 2
NAB_LU_TP0=2.4e-6 NAB_LU_TP1=1.4e-6 python tools/lu_sweep.py 16384 2
