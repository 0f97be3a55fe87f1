set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -4
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_r2_reference.json 2> gpurun_out/bench_r2_reference.err
python bench.py --steps 10 --warmup 3 > gpurun_out/bench_r2_n1.json 2> gpurun_out/bench_r2_n1.err
python -c "
import json; d=json.load(open('gpurun_out/bench_r2_n1.json')); print('bench value %.4g ms %.2f e2e %.4g roofline %.4f share %.3f cpu %.3g launches %d clocks %s' % (d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['frac'], d['roofline']['kernel_share_of_step'], d['cpu_baseline']['value'], d['gpu_launches'], d['clocks']))
r=json.load(open('gpurun_out/bench_r2_reference.json')); print('reference value %.4g cores %d sample %s' % (r['value'], r['cpu_baseline']['cores'], r['config']['sample']))"
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r2.csv python profiles/quick_cfg2.py 1 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_slice_chains -s 60 -c 1 -o gpurun_out/slice_r2 python profiles/quick_cfg2.py 1 2>&1 | tail -1
ncu --set full --clock-control none -k regex:"k_iter_epilogue|k_merge|k_chain_streams|k_iter_advance" -s 240 -c 5 -o gpurun_out/others_r2 python profiles/quick_cfg2.py 1 2>&1 | tail -1
ncu --set full --clock-control none -k regex:"k_radix|k_tree|k_out_degree|k_evidence_stats|k_sort_prep|k_scan_u32" -c 40 -o gpurun_out/finalpass_r2 python profiles/final_pass_once.py 2>&1 | tail -1
timeout 600 compute-sanitizer --tool racecheck --print-limit 5 python profiles/sanitizer_run.py > gpurun_out/sanitizer_racecheck.txt 2>&1; tail -4 gpurun_out/sanitizer_racecheck.txt
timeout 600 compute-sanitizer --tool memcheck --print-limit 5 python profiles/sanitizer_run.py > gpurun_out/sanitizer_memcheck.txt 2>&1; tail -4 gpurun_out/sanitizer_memcheck.txt
