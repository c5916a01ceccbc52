# Regenerates the measurements behind profiles/r01_* on one B200 (run through gpurun from the repo root).
set -x
mkdir -p gpurun_out/final
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/final/bench_reference.json 2> gpurun_out/final/bench_reference.err
python bench.py > gpurun_out/final/bench.json 2> gpurun_out/final/bench.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/final/bench_launches.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu --no-extra > gpurun_out/final/bench_under_ncu.log 2>&1
if [ "$1" = "full" ]; then
ncu --set full --clock-control none --import-source on -k regex:dgemm_tma -s 3 -c 1 -o gpurun_out/final/dgemm_full python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu --no-extra > gpurun_out/final/bench_under_ncu2.log 2>&1
ncu -i gpurun_out/final/dgemm_full.ncu-rep --page raw --csv > gpurun_out/final/dgemm_full_raw.csv 2>/dev/null
ncu -i gpurun_out/final/dgemm_full.ncu-rep --page details > gpurun_out/final/dgemm_full_details.txt 2>/dev/null
fi
for w in chol lu; do ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/final/${w}_launches.csv python tools/prof_factor.py 16384 $w > gpurun_out/final/${w}_prof.log 2>&1; done
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/final/qr_launches.csv python tools/prof_factor.py 65536 qr 4096 > gpurun_out/final/qr_prof.log 2>&1
NAB_LU_TRACE=1 python tools/lu_once.py 2> gpurun_out/final/lu_trace.txt
NAB_CHOL_TRACE=1 python tools/chol_once.py 2> gpurun_out/final/chol_trace.txt
NAB_QR_TRACE=1 python tools/prof_factor.py 65536 qr 4096 2> gpurun_out/final/qr_trace.txt
python tools/factor_timing.py chol,lu,qr > gpurun_out/final/factor_timing.txt 2>&1
