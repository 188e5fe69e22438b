#!/bin/bash
# end-of-round measurement batch (one gpurun call, one GPU); everything lands in gpurun_out/
O=gpurun_out
python bench.py > $O/bench_h.json 2> $O/bench_h.err; tail -c 400 $O/bench_h.err
python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_h_ref.json 2> $O/bench_h_ref.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_r01_h.csv \
    python bench.py --steps 2 --warmup 3 --quick --no-cpu > $O/ncu_bench_h.log 2>&1
ncu --set full --import-source on --clock-control none -k regex:"k_xdelta|k_hzr|k_planes|k_frame|k_scan" --launch-skip 26 --launch-count 16 \
    -o $O/prof_r01_h -f python tools/prof_compress.py xdelta_hzr 592 --dec > $O/ncu_h1.log 2>&1
ncu --set full --import-source on --clock-control none -k regex:"k_fwht|k_hzr_decode|k_hzr_encode$|k_hzr_hist" --launch-skip 12 --launch-count 6 \
    -o $O/prof_r01_h_had -f python tools/prof_compress.py hadamard 592 --dec > $O/ncu_h2.log 2>&1
python tools/config5.py > $O/config5_1gpu_h.jsonl 2> $O/config5_1gpu_h.err
tail -n 2 $O/ncu_h1.log $O/ncu_h2.log; cat $O/bench_h.json | cut -c1-600; cat $O/config5_1gpu_h.jsonl | cut -c1-330
