#!/bin/bash
# Round-2 profiling pass (run on the GPU box through gpurun): launch lists + one `ncu --set full` capture per kernel
# family.  The .ncu-rep files are converted to CSV on the box and deleted (gpurun_out/ is capped at 64 MiB).
mkdir -p gpurun_out
NCU="ncu --set full --clock-control none --import-source on"
cap() {   # cap <name> <source-page? 0|1> <ncu args...> -- <cmd...>
  local name=$1 src=$2; shift 2
  timeout 300 $NCU "$@" -o gpurun_out/$name > gpurun_out/$name.log 2>&1
  if [ -f gpurun_out/$name.ncu-rep ]; then
    ncu -i gpurun_out/$name.ncu-rep --page raw --csv > gpurun_out/$name.raw.csv 2>/dev/null
    [ "$src" = "1" ] && ncu -i gpurun_out/$name.ncu-rep --page source --csv > gpurun_out/$name.src.csv 2>/dev/null
    rm -f gpurun_out/$name.ncu-rep
  fi
}
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches_dcp.csv python profiles/prof_run.py dcp > /dev/null 2>&1
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r02_launches_bench.csv python bench.py --profile --steps 20 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
cap r02_knn32 1 -k regex:knn_kernel -s 3 -c 1 python profiles/prof_run.py knn
cap r02_edge 1 -k regex:edge_ -s 6 -c 6 python profiles/prof_run.py edge
cap r02_emd 1 -k "regex:emd_persistent|emd_final|emd_grad" -s 4 -c 4 python profiles/prof_run.py emd
cap r02_attn 0 -k "regex:softcorr_kernel|edge_gemm|layernorm" -s 7 -c 7 python profiles/prof_run.py attn
cap r02_group 0 -k "regex:fps|furthest|ball|group_points|gather_points" -c 6 python profiles/prof_run.py fps
cap r02_misc 0 -k "regex:svd_head|kabsch|knn_matrix|knn_select|sqnorm|softcorr" -c 8 python profiles/prof_run.py kabsch
cap r02_rpm 0 -k "regex:sinkhorn|weighted_rigid|softcorr_kernel" -s 13 -c 5 python profiles/prof_run.py rpm
cap r02_chamfer 0 -k regex:chamfer -s 4 -c 4 python profiles/prof_run.py chamfer
timeout 400 python bench.py --steps 20 --warmup 3 > gpurun_out/r02_bench.json 2> gpurun_out/r02_bench.err
timeout 400 python -m pytest tests/test_gpu_chamfer.py tests/test_models.py tests/test_gpu_softcorr.py tests/test_gpu_rpm.py tests/test_gpu_emd_svd.py -x -q 2>&1 | tail -12 > gpurun_out/r02_tests.log
du -sh gpurun_out; tail -3 gpurun_out/r02_bench.err; cat gpurun_out/r02_tests.log
