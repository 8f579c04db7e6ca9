#!/bin/bash
mkdir -p gpurun_out
./scripts/gather_probe 0 4096 > gpurun_out/gather_default.txt 2>&1; cat gpurun_out/gather_default.txt
./scripts/gather_probe 32 4096 > gpurun_out/gather_g32.txt 2>&1; cat gpurun_out/gather_g32.txt
./scripts/gather_probe 128 4096 > gpurun_out/gather_g128.txt 2>&1; cat gpurun_out/gather_g128.txt
./scripts/gather_probe 0 64 > gpurun_out/gather_l2resident.txt 2>&1; cat gpurun_out/gather_l2resident.txt
./scripts/gather_probe 0 4096 > /dev/null 2>&1 && ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,lts__t_sectors_srcunit_tex_op_read.sum,l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum --clock-control none --csv --log-file gpurun_out/gather_ncu.csv ./scripts/gather_probe 0 4096 > gpurun_out/gather_ncu.log 2>&1
echo ncu exit $?
python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/gather_ncu.csv')) if len(r)>10]
h=rows[0]; 
for r in rows[1:]:
    d=dict(zip(h,r))
    print(d['ID'], d['Kernel Name'][:30], d['Metric Name'], d['Metric Value'])
PY
