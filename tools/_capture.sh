set -x
B="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-admm --no-extra --no-e2e --no-roofline"
timeout 300 $B > gpurun_out/cap_bench.log 2>&1 || exit 1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2_launches.csv $B > gpurun_out/cap_ncu1.log 2>&1
timeout 120 python tools/gpu_steps.py C3 4 > gpurun_out/cap_steps.log 2>&1 || exit 1
timeout 900 ncu --set full --import-source on --clock-control none -k regex:"blend_(fwd|bwd)" -s 4 -c 2 -f -o gpurun_out/r2_blend python tools/gpu_steps.py C3 4 > gpurun_out/cap_ncu2.log 2>&1
timeout 900 ncu --set full --import-source on --clock-control none -k regex:"preprocess_|radix_scatter|radix_hist|emit_instances|fssim|gather_emit|scan_apply|tile_ranges" -s 40 -c 22 -f -o gpurun_out/r2_other python tools/gpu_steps.py C3 4 > gpurun_out/cap_ncu3.log 2>&1
for c in C1 C2 C3; do timeout 120 python tools/gpu_stages.py $c > /dev/null 2>&1; done
timeout 600 python tools/gpu_truth_compare.py C3 > gpurun_out/cap_truth.log 2>&1
tail -3 gpurun_out/cap_truth.log
ls -la gpurun_out/
