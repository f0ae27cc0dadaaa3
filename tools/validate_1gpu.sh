mkdir -p gpurun_out
rm -f gpurun_out/literal_parity.txt gpurun_out/bvh_build.txt
L=gpurun_out/final.log; : > $L
(timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -8) >> $L 2>&1
(timeout 500 python bench.py > gpurun_out/final_bench.json 2> gpurun_out/final_bench.err)
(timeout 200 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/final_bench_reference.json 2>> gpurun_out/final_bench.err)
python -c "import __graft_entry__ as g; g.smoke()" >> $L 2>&1
timeout 300 python tools/e2e_probe.py >> $L 2>&1
cat $L | tail -25; tail -c 300 gpurun_out/final_bench.json; tail -3 gpurun_out/final_bench.err
echo "== scenes beyond the flat program (17+ primitives: reference-order walk over tables in shared memory)"
timeout 120 python tools/profile_run.py many 1024 1024 8 3
timeout 120 python tools/profile_run.py table 1024 1024 8 3
