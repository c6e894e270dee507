timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/r02bc_tests.log 2>&1; tail -3 gpurun_out/r02bc_tests.log
timeout 900 python bench.py > gpurun_out/r02bc_bench.json 2> gpurun_out/r02bc_bench.err; tail -c 300 gpurun_out/r02bc_bench.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"fused_rows|transform|row_apply|piece_reduce|os_|seg_|cp_|dense_apply|fused_reduce" -c 150 --csv --log-file gpurun_out/r02bc_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-secondary --no-zipf > gpurun_out/r02bc_ncu1.log 2>&1
timeout 600 ncu --set full --import-source on --clock-control none -k regex:fused_rows -s 3 -c 1 -f -o gpurun_out/r02bc_fused_rows python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-secondary --no-zipf > gpurun_out/r02bc_ncu2.log 2>&1
ls -la gpurun_out/r02bc*
