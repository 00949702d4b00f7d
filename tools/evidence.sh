#!/bin/bash
# Round evidence run (on a B200 through gpurun): GPU tests, differential check against the live reference kernels,
# every bench workload + the reference arm, the ncu launch list of the default bench command, and one `ncu --set full`
# report per dominant kernel (each only after the same command has exited 0 without ncu).  Everything lands in
# gpurun_out/; tools/make_profiles.py turns the reports into profiles/rNN_*.txt afterwards.
#   gpurun --timeout 1500 -- 'bash tools/evidence.sh [run|ncu|all] [report names of the ncu part ...]'
# gpurun brings back at most 64 MiB: every ncu capture is limited to its kernel (-k) and one launch (-c).
set -u
O=gpurun_out
mkdir -p $O
WHAT=${1:-all}
if [ "$WHAT" != ncu ]; then
timeout 600 python -m pytest tests -m gpu -q -rs > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?"
timeout 600 python tools/gpu_check.py > $O/gpu_check.log 2>&1; echo "gpu_check rc=$?"
timeout 300 python tools/kitti_check.py > $O/kitti_check.log 2>&1; echo "kitti_check rc=$?"
: > $O/bench_lines.jsonl
timeout 300 python bench.py 2> $O/bench_default.err | tail -1 > $O/bench_default.json; echo "bench default (with secondary + gpu_baseline) rc=$?"
for w in nms_cfg2 nms_cfg5 iou_dense iou_cfg1 iou_cfg4 pib_cfg3 post_cfg2 iou_max_cfg4 roiaware_partA2 roipoint_pointrcnn kitti_eval; do
    timeout 300 python bench.py --workload $w --no-secondary 2> $O/bench_$w.err | tail -1 >> $O/bench_lines.jsonl; echo "bench $w rc=$?"
done
timeout 400 python bench.py --impl reference 2> $O/bench_reference.err | tail -1 >> $O/bench_lines.jsonl; echo "bench reference rc=$?"
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "smoke rc=$?"
LG_LIB_PATH=lidardetection_b200/liblidargeom_timing.so timeout 120 python tools/lz_timing.py > $O/nms_lazy_phase_timing.log 2>&1; echo "lz_timing rc=$?"
# launch list of the default workload (the secondary metrics and the baselines of the default command are left out: their launches
# would fill the 400-launch window before the timed region of the headline is reached)
if timeout 200 python bench.py --steps 2 --warmup 1 --no-secondary --no-cpu-baseline > $O/plain_bench.log 2>&1; then
    timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_nms_cfg2.csv \
        python bench.py --steps 2 --warmup 1 --no-secondary --no-cpu-baseline > $O/ncu_launch.log 2>&1; echo "launch list rc=$?"
fi
fi
if [ "$WHAT" != run ]; then
for tk in nms64:nms_lazy_kernel nms_cfg5:nms_lazy_kernel nms_full:nms_mask_kernel iou_dense16k:iou_strip_kernel:2 iou_cfg4:iou_sweep_kernel:1 iou_cfg4:iou_pairs_kernel:1:iou_cfg4_pairs iou_cfg1:iou_flat_kernel pib4096:pib_grid_kernel kitti:kitti_pair_kernel roiaware:roiaware_collect_pool_kernel roipoint:roipoint_pool_kernel; do
    IFS=: read -r t k c nm <<< "$tk"; c=${c:-1}; nm=${nm:-$t}
    if [ $# -gt 1 ] && [[ " ${*:2} " != *" $nm "* ]]; then continue; fi   # `evidence.sh ncu name ...`: only these reports (64 MiB per call)   # target : kernel regex : launches to capture (the strip kernel is launched in two builds) : report name
    if timeout 120 python tools/prof_target.py $t 3 > $O/plain_$nm.log 2>&1; then
        timeout 500 ncu --set full --clock-control none --import-source on -k regex:$k -c $c -f -o $O/final_$nm \
            python tools/prof_target.py $t 2 > $O/ncu_$nm.log 2>&1; echo "ncu $nm rc=$?"
    else
        echo "plain $t failed"
    fi
done
fi
ls -la $O/*.ncu-rep | tail -8
du -sm $O
