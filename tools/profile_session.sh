#!/bin/bash
# One GPU session that produces everything profiles/ cites for the current build (run under gpurun):
#   launch list of a bench step, ncu --set full captures of the wavefront kernel on the three scene
#   classes (SMEM-resident / multi-light / deep mesh), and of the finish + denoise kernels.
tag=${1:-r01i}
out=gpurun_out
NCU="ncu --set full --clock-control none --import-source on"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/launches_$tag.csv python bench.py --steps 2 --warmup 1 > $out/bench_under_ncu.log 2>&1
timeout 200 $NCU -k regex:k_wavefront2 -s 1 -c 1 -f -o $out/prof_${tag}_cornell python tools/profile_run.py cornell 1024 1024 32 1 > /dev/null 2>&1
timeout 200 $NCU -k regex:k_wavefront2 -s 1 -c 1 -f -o $out/prof_${tag}_veach python tools/profile_run.py veach 1024 1024 16 1 > /dev/null 2>&1
timeout 200 $NCU -k regex:k_wavefront2 -s 1 -c 1 -f -o $out/prof_${tag}_ajax python tools/profile_run.py ajax 1024 1024 16 1 > /dev/null 2>&1
timeout 200 $NCU -k regex:"k_finish|k_nlm" -c 3 -f -o $out/prof_${tag}_finish python tools/finish_bench.py cornell 1024 1024 > /dev/null 2>&1
ls -la $out/prof_${tag}_*.ncu-rep
