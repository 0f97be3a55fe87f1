# end-of-round evidence on one B200: the ncu launch list of the contract command itself, and the sanitizer logs
mkdir -p gpurun_out/r2f
O=gpurun_out/r2f
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/launches_bench.csv python bench.py --steps 2 --warmup 1 > $O/bench_under_ncu.json 2> $O/bench_under_ncu.err
python profiles/launch_summary.py $O/launches_bench.csv > $O/launches_bench_summary.txt 2>&1; head -c 2000000 $O/launches_bench.csv > $O/launches_bench_head.csv; rm -f $O/launches_bench.csv; head -24 $O/launches_bench_summary.txt
timeout 900 compute-sanitizer --tool racecheck --print-limit 5 python profiles/sanitizer_run.py > $O/sanitizer_racecheck.txt 2>&1; tail -4 $O/sanitizer_racecheck.txt
timeout 900 compute-sanitizer --tool memcheck --print-limit 5 python profiles/sanitizer_run.py > $O/sanitizer_memcheck.txt 2>&1; tail -4 $O/sanitizer_memcheck.txt
du -sh $O
