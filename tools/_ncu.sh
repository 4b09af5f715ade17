W="python tools/ncu_workload.py"
O=gpurun_out
Q="dram__bytes_read.sum|dram__bytes_write.sum|gpu__time_duration.sum|sm__pipe_tensor_cycles_active|sm__warps_active.avg.pct_of_peak|launch__registers_per_thread|launch__grid_size|launch__block_size|sm__throughput.avg.pct|gpu__dram_throughput|lts__t_bytes.sum|l1tex__data_bank_conflicts|smsp__inst_executed.sum |sm__inst_executed_pipe_tensor|launch__occupancy_limit|lts__t_sector_hit_rate|dram__throughput"
tocsv() { ncu -i $O/$1.ncu-rep --page raw --csv > $O/$1_raw.csv 2>/dev/null; rm -f $O/$1.ncu-rep; }
# launch lists (cheap)
$W cfg2 > $O/r2g_plain_cfg2.log 2>&1 && timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/r2g_cfg2_launches.csv $W cfg2 > $O/r2g_ncu1.log 2>&1
$W cfg4 > $O/r2g_plain_cfg4.log 2>&1 && timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/r2g_cfg4_launches.csv $W cfg4 > $O/r2g_ncu4.log 2>&1
$W probe > $O/r2g_plain_probe.log 2>&1 && timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/r2g_probe_launches.csv $W probe > $O/r2g_ncu7.log 2>&1
# full sets: a few launches of every kernel kind
K="regex:gemm_tc|attention_tc|reduce_ln|head_decode|crop_resize|box_overlay|gather_template"
timeout 500 ncu --set full --clock-control none --cache-control none -k "$K" -s 140 -c 16 -f -o $O/r2g_cfg2_warm $W cfg2 > $O/r2g_ncu3.log 2>&1; tocsv r2g_cfg2_warm
timeout 500 ncu --set full --clock-control none -k "$K" -s 140 -c 16 -f -o $O/r2g_cfg2_cold $W cfg2 > $O/r2g_ncu2.log 2>&1; tocsv r2g_cfg2_cold
K4="regex:gemm_as|gemm_tc|attention_tc|reduce_ln"
timeout 500 ncu --set full --clock-control none --cache-control none -k "$K4" -s 130 -c 14 -f -o $O/r2g_cfg4_warm $W cfg4 > $O/r2g_ncu5.log 2>&1; tocsv r2g_cfg4_warm
$W pixels > $O/r2g_plain_pixels.log 2>&1 && timeout 400 ncu --set full --clock-control none -k regex:"nv12_to_rgb|yuy2_to_rgb|resize_rgb" -f -o $O/r2g_pixels $W pixels > $O/r2g_ncu6.log 2>&1; tocsv r2g_pixels
# one small report with sources for the dominant kernel
timeout 300 ncu --set full --clock-control none --cache-control none --import-source on -k regex:gemm_tc -s 60 -c 2 -f -o $O/r2g_gemm_tc_src $W cfg2 > $O/r2g_ncu8.log 2>&1
ls -la $O/ | head -40; du -sh $O
