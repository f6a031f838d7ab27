# scratch: the command list of the last gpurun session (development aid)
cd /root/repo
timeout 600 ncu --set full --import-source on --clock-control none -k regex:window_attention_tc -s 1 -c 1 -f -o gpurun_out/att_tc python tools/att_one.py 2>&1 | tail -1
timeout 600 ncu --set full --import-source on --clock-control none -k regex:cva_offsets_reg -s 2 -c 1 -f -o gpurun_out/cva_off python tools/cva_one.py 2>&1 | tail -1
