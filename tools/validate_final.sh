# round-end validation on one B200: ncu capture of the bench kernel, its constants into algo_bytes.json, the whole
# GPU suite, both bench arms, the launch list, smoke()
mkdir -p gpurun_out
rm -f gpurun_out/literal_parity.txt gpurun_out/bvh_build.txt
L=gpurun_out/final.log; : > $L
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_wavefront2 -s 1 -c 1 -f -o gpurun_out/prof_r02_cornell_lq python tools/profile_run.py cornell 1024 1024 32 1 >> $L 2>&1
python tools/ncu_constants.py gpurun_out/prof_r02_cornell_lq.ncu-rep cornell 33554432 "profiles/r02_k_wavefront2_cornell.txt (ncu --set full, one launch of 33554432 samples; traversal share = ray_mesh + ray_aabb + ray_tri + prim_test + trace_closest + their inlined vector math, % of samples)" >> $L 2>&1
cp tools/algo_bytes.json gpurun_out/algo_bytes.json
(timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -8) >> $L 2>&1
(timeout 500 python bench.py > gpurun_out/final_bench.json 2> gpurun_out/final_bench.err)
(timeout 200 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/final_bench_reference.json 2>> gpurun_out/final_bench.err)
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches.csv python bench.py --steps 2 --warmup 1 --no-configs > gpurun_out/final_bench_under_ncu.log 2>&1
python -c "import __graft_entry__ as g; g.smoke()" >> $L 2>&1
cat $L | grep -v "^==PROF\|^==WARN" | tail -25; tail -c 300 gpurun_out/final_bench.json; tail -3 gpurun_out/final_bench.err
