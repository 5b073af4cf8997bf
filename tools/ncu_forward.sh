# ncu evidence for one RetinaFace-R50 forward over 64 x 1080p frames (run on ONE GPU after the same program has exited 0 without ncu):
#  * launch list with time, DRAM bytes, L2 bytes, tensor-pipe activity of every kernel  -> gpurun_out/r02_fwd_launches_v4.csv
#  * --set full capture of the fused-tail kernels of the second forward                 -> gpurun_out/r02_conv_tail_full_v2.ncu-rep
set -x
python tools/forward_once.py > gpurun_out/fwd_plain.log 2>&1 || exit 1
cat gpurun_out/fwd_plain.log
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active --clock-control none --csv --log-file gpurun_out/r02_fwd_launches_v4.csv python tools/forward_once.py > gpurun_out/ncu_list.log 2>&1
echo list_rc=$?
ncu --set full --clock-control none --import-source on -k regex:conv_tail -s 6 -c 6 -o gpurun_out/r02_conv_tail_full_v2 python tools/forward_once.py > gpurun_out/ncu_full.log 2>&1
echo full_rc=$?
ncu -i gpurun_out/r02_conv_tail_full_v2.ncu-rep --page raw --csv > gpurun_out/r02_conv_tail_full_v2_raw.csv 2>/dev/null
ls -la gpurun_out/*.ncu-rep | tail -2
