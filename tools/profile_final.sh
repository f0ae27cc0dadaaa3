mkdir -p gpurun_out
rm -f gpurun_out/literal_parity.txt gpurun_out/bvh_build.txt
L=gpurun_out/prof.log; : > $L
(timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -8) >> $L 2>&1
(timeout 500 python bench.py > gpurun_out/prof_bench.json 2> gpurun_out/prof_bench.err)
(timeout 200 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/prof_bench_reference.json 2>> gpurun_out/prof_bench.err)
prof() { timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_wavefront2 -s 1 -c 1 -o gpurun_out/prof_r02_$1 python tools/profile_run.py $1 $2 $3 $4 1 >> $L 2>&1; }
prof cornell 1024 1024 32
prof veach 1920 1080 16
prof ajax 1024 1024 16
prof env 2048 2048 16
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches.csv python bench.py --steps 2 --warmup 1 --no-configs > gpurun_out/prof_bench_under_ncu.log 2>&1
python -c "import __graft_entry__ as g; g.smoke()" >> $L 2>&1
cat $L | grep -v "^==PROF\|^==WARN" | tail -30; tail -c 300 gpurun_out/prof_bench.json; tail -3 gpurun_out/prof_bench.err
