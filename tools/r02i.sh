mkdir -p gpurun_out
rm -f gpurun_out/literal_parity.txt gpurun_out/bvh_build.txt
(timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -8) > gpurun_out/r02i_tests.log 2>&1
(timeout 500 python bench.py > gpurun_out/r02i_bench.json 2> gpurun_out/r02i_bench.err)
(timeout 200 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02i_bench_reference.json 2>> gpurun_out/r02i_bench.err)
timeout 1500 bash tools/racecheck.sh gpurun_out/r02i_racecheck.txt > /dev/null 2>&1
cat gpurun_out/r02i_tests.log; tail -c 400 gpurun_out/r02i_bench.json; tail -3 gpurun_out/r02i_bench.err; cat gpurun_out/r02i_racecheck.txt | head -60
