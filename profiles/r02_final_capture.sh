#!/bin/bash
# Round-2 closing profile pass (run on the GPU box through gpurun).  Same conventions as r02_capture.sh: the
# `ncu --set full` details pages go to gpurun_out/r02_<family>.log (text), launch lists to CSV; nothing large is kept.
mkdir -p gpurun_out
NCU="ncu --set full --clock-control none"
cap() {   # cap <name> <ncu filter args...> -- python ...   (details page -> gpurun_out/<name>.log)
  local name=$1; shift
  timeout 300 $NCU "$@" > gpurun_out/$name.log 2>&1
}
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 500 --csv --log-file gpurun_out/r02_launches_dcp.csv python profiles/prof_run.py dcp > /dev/null 2>&1
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r02_launches_bench.csv python bench.py --profile --steps 20 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
cap r02_knn32 -k regex:knn_kernel -s 3 -c 1 python profiles/prof_run.py knn
cap r02_edge -k regex:edge_ -s 6 -c 6 python profiles/prof_run.py edge
cap r02_attn -k "regex:softcorr_kernel|edge_gemm|attn_" -s 9 -c 9 python profiles/prof_run.py attn
cap r02_knnstream -k regex:knn_stream -s 1 -c 1 python profiles/prof_run.py knnstream
timeout 120 python profiles/attention_bounds.py > gpurun_out/r02_attention_bounds.txt 2>&1
timeout 400 python bench.py --steps 20 --warmup 3 > gpurun_out/r02_bench.json 2> gpurun_out/r02_bench.err
timeout 400 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/r02_bench_reference.json 2> gpurun_out/r02_bench_reference.err
timeout 300 python profiles/time_models.py > gpurun_out/r02_time_models_final.jsonl 2>/dev/null
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 > gpurun_out/r02_tests.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/r02_smoke.log 2>&1
du -sh gpurun_out; tail -2 gpurun_out/r02_bench.err; cat gpurun_out/r02_tests.log gpurun_out/r02_smoke.log; cat gpurun_out/r02_attention_bounds.txt | tail -8
