#!/bin/bash
# End-of-round pass (run under gpurun): the whole GPU suite, the bench (both arms), the captures of the kernels that changed late in the round.
set -u
O=gpurun_out
mkdir -p $O
python -m pytest tests -m gpu -x -q > $O/r2_gputest_final.log 2>&1; echo "rc=$?" >> $O/r2_gputest_final.log; tail -3 $O/r2_gputest_final.log
python bench.py > $O/r2_bench_final.json 2> $O/r2_bench_final.err; echo "bench rc=$?"
python bench.py --impl reference > $O/r2_bench_reference.json 2>> $O/r2_bench_final.err; echo "reference rc=$?"
python -c "import __graft_entry__ as g; g.smoke()" > $O/r2_smoke.log 2>&1; echo "smoke rc=$?"
NCU="ncu --set full --clock-control none --import-source on -c 1"
$NCU -k regex:graph_planar_tma -s 1 -o $O/prof_gvoice_planar python tools/prof_one.py gvoice 1 > /dev/null 2>&1
XV_N=524288 XV_STEADY=1 $NCU -k regex:k_xvoice_mix2 -s 2 -o $O/prof_xvoice_mix2_shard python tools/prof_one.py xvoice 2 > /dev/null 2>&1
SUMMARY_DIR=$O/summaries python tools/summarize_ncu.py $O/prof_gvoice_planar.ncu-rep $O/prof_xvoice_mix2_shard.ncu-rep
rm -f $O/prof_gvoice_planar.ncu-rep $O/prof_xvoice_mix2_shard.ncu-rep
head -c 600 $O/r2_bench_final.json; echo
