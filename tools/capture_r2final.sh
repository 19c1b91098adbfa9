# Round-end capture on one GPU (gpurun): smoke, all single-GPU bench lines, step traces, ncu of the loss step, launch list.
set -x
cd $GRAFT_REPO_ROOT
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2g_smoke.log 2>&1; tail -2 gpurun_out/r2g_smoke.log
sed -i 's/r2f_/r2g_/g' tools/bench_sweep_1gpu.sh
bash tools/bench_sweep_1gpu.sh
python tools/loss_once.py 16 nchw > gpurun_out/r2g_plain.log 2>&1 && python tools/loss_once.py 16 nhwc >> gpurun_out/r2g_plain.log 2>&1 || exit 1
ncu --set full --clock-control none --import-source on -k regex:"prep_step|iou_match|match_score|select_gmm|bulk_focal|positive_list" --launch-skip 12 -c 6 -f -o gpurun_out/r2g_loss python tools/loss_once.py 16 nchw > gpurun_out/r2g_ncu_loss.log 2>&1
python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-side --no-graph > gpurun_out/r2g_bench_nograph.json 2> gpurun_out/r2g_bench_nograph.err && ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2g_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-side --no-graph > gpurun_out/r2g_ncu_bench.log 2>&1
ls -la gpurun_out/r2g_*
