# Round-end capture on one GPU (gpurun): GPU test suite, smoke, the loss path's single-GPU bench lines, step traces, ncu of
# the loss step, launch list.  (The post-processing lines / captures are tools/bench_sweep_1gpu.sh and tools/capture_r2z.sh.)
set -x
cd $GRAFT_REPO_ROOT
python -m pytest tests -x -q -m gpu > gpurun_out/r2h_tests.log 2>&1; tail -2 gpurun_out/r2h_tests.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2h_smoke.log 2>&1; tail -2 gpurun_out/r2h_smoke.log
python bench.py > gpurun_out/r2h_bench.json 2> gpurun_out/r2h_bench.err
python bench.py --layout nhwc --no-post --no-cpu-baseline > gpurun_out/r2h_bench_nhwc.json 2>> gpurun_out/r2h_bench.err
for c in C1 C3 C5; do python bench.py --config $c --no-post --no-side --no-cpu-baseline > gpurun_out/r2h_bench_$c.json 2>> gpurun_out/r2h_bench.err; done
python bench.py --images-per-gpu 2 --no-post --no-side --no-cpu-baseline > gpurun_out/r2h_bench_2img.json 2>> gpurun_out/r2h_bench.err
(python tools/step_trace.py --images 16; python tools/step_trace.py --images 2; python tools/step_trace.py --post --images 64; python tools/step_trace.py --post --images 8) 2>/dev/null > gpurun_out/r2h_step_trace.txt
python tools/loss_once.py 16 nchw > gpurun_out/r2h_plain.log 2>&1 || exit 1
ncu --set full --clock-control none --import-source on -k regex:"prep_step|iou_match|match_score|select_gmm|bulk_focal|positive_list" --launch-skip 12 -c 6 -f -o gpurun_out/r2h_loss python tools/loss_once.py 16 nchw > gpurun_out/r2h_ncu_loss.log 2>&1
python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-side --no-graph > gpurun_out/r2h_bench_nograph.json 2> gpurun_out/r2h_bench_nograph.err && ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2h_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-side --no-graph > gpurun_out/r2h_ncu_bench.log 2>&1
ls gpurun_out/r2h_*
