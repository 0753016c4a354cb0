#!/bin/bash
mkdir -p gpurun_out
timeout 900 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_sector_hit_rate.pct,lts__t_sectors_srcunit_tex_op_read.sum,lts__t_sectors_op_read.sum,lts__t_sectors_op_write.sum,l1tex__m_xbar2l1tex_read_bytes.sum --clock-control none -k regex:"gemm_tcgen05|wgrad_halo" --csv --log-file gpurun_out/r02_conv3_dram.csv python tools/ncu_target.py conv3 > gpurun_out/r02_conv3_dram.log 2>&1; echo "ncu rc=$?"; tail -2 gpurun_out/r02_conv3_dram.log
