set -x
python tools/forward_once.py > gpurun_out/fwd_plain.log 2>&1 || exit 1
cat gpurun_out/fwd_plain.log
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active --clock-control none --csv --log-file gpurun_out/r02_fwd_launches_v3.csv python tools/forward_once.py > gpurun_out/ncu_list.log 2>&1
echo list_rc=$?
ncu --set full --clock-control none --import-source on -k regex:conv_tail -s 3 -c 3 -o gpurun_out/r02_conv_tail_full python tools/forward_once.py > gpurun_out/ncu_full.log 2>&1
echo full_rc=$?
ls -la gpurun_out/*.ncu-rep
