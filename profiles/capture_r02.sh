#!/bin/bash
# ncu captures of round 2 (run on the GPU box through gpurun; outputs under gpurun_out/, summarised into profiles/
# with profiles/summarize.py and profiles/launches_summary.py).  Numbers measured under ncu are never bench values.
# Kernel names are matched in their mangled form: StaticSpec<7, 0x2222222, true, WIDE, RotX, RotZ, true> =
# StaticSpecILi7ELy35791394ELb1ELb<WIDE>E..., PatternSpec<identity, identity, WIDE> = PatternSpecILi1ELi1ELb<WIDE>E.
B="python bench.py --steps 1 --warmup 0 --no-cpu-baseline --no-other-configs --no-pipelined --spot-check 0"
NCU="ncu --clock-control none --kernel-name-base mangled"
if [ "$1" != "full-only" ]; then
# every launch of the bench command with its device time and DRAM traffic
$NCU --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum -c 1500 --csv --log-file gpurun_out/r02_launches.csv $B > gpurun_out/r02_launches.out 2>&1
fi
# --set full of the dominant kernel: generation 1 (throughput flavour, ~65 k problems) and generation 70 (wide flavour)
$NCU --set full --import-source on --kernel-name "regex:memetic_generation_kernel.*StaticSpecILi7ELy35791394ELb1ELb0E" --launch-skip 1 --launch-count 1 -f -o gpurun_out/r02_panda_bulk $B > gpurun_out/r02_panda_bulk.out 2>&1
$NCU --set full --import-source on --kernel-name "regex:memetic_generation_kernel.*StaticSpecILi7ELy35791394ELb1ELb1E" --launch-skip 70 --launch-count 1 -f -o gpurun_out/r02_panda_wide $B > gpurun_out/r02_panda_wide.out 2>&1
# Fetch-256 (BASELINE configs[3]): generation 1 and a wide generation; UR5 GD-only (configs[2])
$NCU --set full --import-source on --kernel-name "regex:memetic_generation_kernel.*PatternSpecILi1ELi1ELb0E" --launch-skip 1 --launch-count 1 -f -o gpurun_out/r02_fetch_bulk python profiles/fetch_probe.py 65536 > gpurun_out/r02_fetch_bulk.out 2>&1
$NCU --set full --import-source on --kernel-name "regex:memetic_generation_kernel.*PatternSpecILi1ELi1ELb1E" --launch-skip 40 --launch-count 1 -f -o gpurun_out/r02_fetch_wide python profiles/fetch_probe.py 65536 > gpurun_out/r02_fetch_wide.out 2>&1
if true; then
$NCU --set full --import-source on --kernel-name "regex:gd_local_kernel" --launch-skip 0 --launch-count 1 -f -o gpurun_out/r02_ur5_gd_local python profiles/fetch_probe.py 262144 ur5 > gpurun_out/r02_ur5_gd_local.out 2>&1
fi
# summaries made on the box (the reports together exceed what gpurun copies back): raw-page summary, per-segment
# instruction / stall profile of the source page; only the two Panda reports travel
for name in r02_panda_bulk r02_panda_wide r02_fetch_bulk r02_fetch_wide r02_ur5_gd_local; do
  [ -f gpurun_out/$name.ncu-rep ] || continue
  python profiles/summarize.py gpurun_out/$name.ncu-rep > gpurun_out/$name.txt 2>&1
  ncu -i gpurun_out/$name.ncu-rep --page source --csv > gpurun_out/$name.source.csv 2>/dev/null
  python profiles/srcpage.py gpurun_out/$name.source.csv > gpurun_out/${name}_segments.txt 2>&1
  rm -f gpurun_out/$name.source.csv
done
rm -f gpurun_out/r02_fetch_bulk.ncu-rep gpurun_out/r02_fetch_wide.ncu-rep gpurun_out/r02_ur5_gd_local.ncu-rep
ls -la gpurun_out/
