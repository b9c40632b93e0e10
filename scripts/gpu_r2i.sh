#!/bin/bash
# round 2, call I: ncu --set full of the wide-palette kernel on the small R-MAT (first sweep and a late sweep), launch list
mkdir -p gpurun_out
python bench.py --workload c4small --ncol 1024 --quick --steps 2 --warmup 1 > gpurun_out/r2i_plain.json 2>&1 || { tail -5 gpurun_out/r2i_plain.json; exit 1; }
cat gpurun_out/r2i_plain.json | cut -c1-400
ncu --set full --clock-control none --import-source on -k regex:wide_sweep -s 1 -c 1 -o gpurun_out/r2i_wide_first python bench.py --workload c4small --ncol 1024 --quick --steps 2 --warmup 1 > gpurun_out/r2i_ncu1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:wide_sweep -s 10 -c 1 -o gpurun_out/r2i_wide_late python bench.py --workload c4small --ncol 1024 --quick --steps 2 --warmup 1 > gpurun_out/r2i_ncu2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:binned_sweep -s 1 -c 1 -o gpurun_out/r2i_binned_first python bench.py --workload c4small --ncol 512 --quick --steps 2 --warmup 1 > gpurun_out/r2i_ncu3.log 2>&1
ls -la gpurun_out/r2i*.ncu-rep
