import csv,sys,subprocess
want=['Kernel Name','Grid Size','gpu__time_duration.sum','dram__bytes_read.sum','dram__bytes_write.sum','dram__throughput.avg.pct_of_peak_sustained_elapsed','gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed','sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed','TPC.TriageCompute.sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed','sm__inst_executed_pipe_tmem.avg.pct_of_peak_sustained_active','sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed','sm__warps_active.avg.pct_of_peak_sustained_active','launch__registers_per_thread','sm__throughput.avg.pct_of_peak_sustained_elapsed','lts__t_bytes.sum','lts__throughput.avg.pct_of_peak_sustained_elapsed','l1tex__throughput.avg.pct_of_peak_sustained_elapsed','dram__bytes_read.sum.per_second','sm__cycles_elapsed.avg.per_second','lts__t_sectors_srcunit_tex_op_read.sum','lts__t_sector_hit_rate.pct','launch__shared_mem_per_block_dynamic']
for f in sys.argv[1:]:
    out=subprocess.run(['ncu','-i',f,'--page','raw','--csv'],capture_output=True,text=True).stdout
    rows=list(csv.reader(out.splitlines()))
    hdr=rows[0]; units=rows[1]
    print('===',f)
    for r in rows[2:]:
        print('--- launch')
        for w in want:
            for i,h in enumerate(hdr):
                if h==w: print(f'   {w:85s} {r[i]} {units[i]}')
