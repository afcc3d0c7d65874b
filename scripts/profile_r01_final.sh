#!/bin/bash
# final round-1 pass: tests, smoke, bench (both arms), then the ncu launch list and one full capture of the top kernel
python -m pytest tests -m gpu -x -q > gpurun_out/f_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/f_pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/f_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/f_smoke.log
python bench.py > gpurun_out/f_bench_n1.json 2> gpurun_out/f_bench_n1.err; echo "bench rc=$?"
python bench.py --impl reference > gpurun_out/f_ref_n1.json 2> gpurun_out/f_ref_n1.err; echo "ref rc=$?"
CMD="python bench.py --views 8 --steps 1 --warmup 1 --no-e2e --no-cpu-baseline"
$CMD > gpurun_out/f_plain.log 2>&1 || { echo "plain run failed"; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/f_launches.csv $CMD > gpurun_out/f_ncu_list.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_sweep -s 8 -c 1 -f -o gpurun_out/f_prof_sweep $CMD > gpurun_out/f_ncu_sweep.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_fuse_view -s 4 -c 1 -f -o gpurun_out/f_prof_fuse_view $CMD > gpurun_out/f_ncu_fuse.log 2>&1
ls gpurun_out/f_*
