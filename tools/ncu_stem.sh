# full capture (with source counters) of the stem conv of the second forward: launch index 1 of conv_tma kernels after skipping the first forward
python tools/forward_once.py > gpurun_out/fwd_plain.log 2>&1 || exit 1
N=$(grep -o "launches per forward: [0-9]*" gpurun_out/fwd_plain.log | grep -o "[0-9]*$")
echo "launches per forward $N"
ncu --set full --clock-control none --import-source on -k regex:conv_tma -s 63 -c 1 -o gpurun_out/r02_stem_pool_full python tools/forward_once.py > gpurun_out/ncu_full.log 2>&1
echo full_rc=$?
ncu -i gpurun_out/r02_stem_pool_full.ncu-rep --page raw --csv > gpurun_out/r02_stem_pool_raw.csv 2>/dev/null
ncu -i gpurun_out/r02_stem_pool_full.ncu-rep --page source --csv > gpurun_out/r02_stem_pool_source.csv 2>/dev/null
ls -la gpurun_out/r02_stem_pool*
