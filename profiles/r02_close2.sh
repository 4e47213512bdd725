#!/bin/bash
# Round-2, second closing pass (thread-per-row kNN kernel): launch list of the bench command, one `ncu --set full`
# details page of knn_duo_kernel, the bench line, the GPU test-suite and smoke().  Run on the GPU box through gpurun.
mkdir -p gpurun_out
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r02b_launches_bench.csv python bench.py --profile --steps 20 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
timeout 200 ncu --set full --clock-control none -k regex:knn_duo -s 2 -c 1 python profiles/prof_knn_tpr.py 32 0 > gpurun_out/r02b_knn_duo.log 2>&1
timeout 400 python bench.py --steps 20 --warmup 3 > gpurun_out/r02b_bench.json 2> gpurun_out/r02b_bench.err
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 > gpurun_out/r02b_tests.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/r02b_smoke.log 2>&1
tail -2 gpurun_out/r02b_bench.err; cat gpurun_out/r02b_tests.log gpurun_out/r02b_smoke.log | tail -6; head -c 400 gpurun_out/r02b_bench.json
