#!/bin/bash
# Round-end evidence on one B200: parity suite, smoke, both bench arms, ncu launch list + full capture of the
# slice kernel, all-config sweep, log Z validation over 10 seeds, device timeline.  Outputs -> gpurun_out/*_final*
O=gpurun_out
if [ -z "$SKIP_TESTS" ]; then
python -m pytest tests -m gpu -q 2>&1 | tail -3 > $O/pytest_gpu_final.txt; cat $O/pytest_gpu_final.txt
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1 > $O/smoke_final.txt; cat $O/smoke_final.txt
fi
python bench.py > $O/bench_final_n1.json 2> $O/bench_final_n1.err; cut -c1-200 $O/bench_final_n1.json
python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_final_reference.json 2> /dev/null; cut -c1-200 $O/bench_final_reference.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file $O/launches_final.csv python bench.py --steps 1 --warmup 3 > $O/launches_final_bench.log 2>&1
python profiles/launch_summary.py $O/launches_final.csv > $O/launches_final_summary.txt; cat $O/launches_final_summary.txt | head -12
ncu --set full --clock-control none --import-source on -k regex:k_slice_chains -s 60 -c 1 -o $O/slice_final -f python profiles/quick_cfg2.py 1 > $O/slice_final.log 2>&1
python profiles/ncu_summary.py $O/slice_final.ncu-rep > $O/slice_final.txt; cat $O/slice_final.txt
python profiles/ncu_lines.py $O/slice_final.ncu-rep > $O/slice_final_lines.txt 2>/dev/null
ncu --set full --clock-control none -k regex:"k_iter_epilogue_grid|k_chain_streams|k_merge_sort_tiles|k_merge_rank|k_merge_scatter" -s 300 -c 6 -o $O/others_final -f python profiles/quick_cfg2.py 1 > $O/others_final.log 2>&1
python profiles/ncu_summary.py $O/others_final.ncu-rep > $O/others_final.txt
python profiles/config_sweep.py > $O/config_sweep_final.txt 2> /dev/null; cut -c1-200 $O/config_sweep_final.txt
python profiles/validate_logz.py > $O/validate_logz_final.txt 2>&1; cat $O/validate_logz_final.txt
NSB200_TRACE=1 python profiles/quick_cfg2.py 1 2>&1 | grep "^trace" | tail -18 > $O/timeline_final.txt; cat $O/timeline_final.txt
rm -f $O/*.ncu-rep  # the summaries above are what is kept (gpurun copies back at most 64 MiB)
