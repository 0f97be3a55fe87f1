# round-2 measurement batch (one B200): contract bench lines, launch list, ncu captures reduced to text on the box
# (gpurun brings back at most 64 MiB), sanitizer runs.  usage: bash profiles/run_r2_profiles.sh [sanitize]
mkdir -p gpurun_out/r2
O=gpurun_out/r2
python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_r2_reference.json 2> $O/bench_r2_reference.err
python bench.py --steps 10 --warmup 3 > $O/bench_r2_n1.json 2> $O/bench_r2_n1.err
python -c "
import json; d=json.load(open('$O/bench_r2_n1.json')); print('bench value %.4g ms %.2f e2e %.4g roofline %.4f share %.3f cpu %.3g launches %d clocks %s' % (d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['frac'], d['roofline']['kernel_share_of_step'], d['cpu_baseline']['value'], d['gpu_launches'], d['clocks']))
r=json.load(open('$O/bench_r2_reference.json')); print('reference value %.4g cores %d sample %s' % (r['value'], r['cpu_baseline']['cores'], r['config']['sample']))"
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/launches_r2.csv python profiles/quick_cfg2.py 1 > /dev/null 2>&1
python profiles/launch_summary.py $O/launches_r2.csv > $O/launches_r2_summary.txt 2>&1; head -c 3000000 $O/launches_r2.csv > $O/launches_r2_head.csv; rm -f $O/launches_r2.csv; tail -15 $O/launches_r2_summary.txt
ncu --set full --clock-control none --import-source on -k regex:k_slice_chains -s 60 -c 1 -o /tmp/slice_r2 python profiles/quick_cfg2.py 1 > /dev/null 2>&1
python profiles/ncu_summary.py /tmp/slice_r2.ncu-rep > $O/slice_r2.txt 2>&1; python profiles/ncu_lines.py /tmp/slice_r2.ncu-rep 970000 40 > $O/slice_r2_lines.txt 2>&1; cat $O/slice_r2.txt | head -30
ncu --set full --clock-control none -k regex:"k_iter_epilogue|k_merge|k_chain_streams|k_iter_advance|k_iter_prologue|k_append_live" -s 480 -c 8 -o /tmp/others_r2 python profiles/quick_cfg2.py 1 > /dev/null 2>&1
python profiles/ncu_summary.py /tmp/others_r2.ncu-rep > $O/others_r2.txt 2>&1
ncu --set full --clock-control none -k regex:"k_radix|k_tree|k_out_degree|k_evidence_stats|k_sort_prep|k_scan_u32" -c 40 -o /tmp/finalpass_r2 python profiles/final_pass_once.py > /dev/null 2>&1
python profiles/ncu_summary.py /tmp/finalpass_r2.ncu-rep > $O/finalpass_r2.txt 2>&1; grep -E "kernel:|gpu__time_duration|dram__bytes" $O/finalpass_r2.txt | head -60
if [ "$1" = "sanitize" ]; then
  timeout 600 compute-sanitizer --tool racecheck --print-limit 5 python profiles/sanitizer_run.py > $O/sanitizer_racecheck.txt 2>&1; tail -3 $O/sanitizer_racecheck.txt
  timeout 600 compute-sanitizer --tool memcheck --print-limit 5 python profiles/sanitizer_run.py > $O/sanitizer_memcheck.txt 2>&1; tail -3 $O/sanitizer_memcheck.txt
fi
du -sh $O
