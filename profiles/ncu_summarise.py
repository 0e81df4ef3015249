import csv, subprocess, io, sys
def raw(rep):
    out = subprocess.run(["ncu","-i",rep,"--page","raw","--csv"],capture_output=True,text=True).stdout
    rows=list(csv.reader(io.StringIO(out))); return rows[0], rows[1], rows[2:]
keep = ['gpu__time_duration.sum','dram__bytes_read.sum','dram__bytes_write.sum','gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed','sm__warps_active.avg.pct_of_peak_sustained_active','launch__registers_per_thread','launch__occupancy_limit_registers','sm__throughput.avg.pct_of_peak_sustained_elapsed','sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active','lts__t_sector_hit_rate.pct','l1tex__t_sector_hit_rate.pct','smsp__inst_executed.sum','launch__grid_size','launch__block_size','l1tex__t_bytes_pipe_lsu_mem_global_op_ld.sum','lts__t_bytes.sum']
title, outp, reps = sys.argv[1], sys.argv[2], sys.argv[3:]
with open(outp,'w') as f:
    f.write(f"# {title}\n\n`ncu --set full --clock-control none --import-source on`, one GPU, under gpurun; numbers are per launch.\n\n")
    for rep in reps:
        hdr, units, rows = raw(rep)
        for r in rows:
            f.write(f"## {r[hdr.index('Kernel Name')]}\n\n| metric | value | unit |\n|---|---|---|\n")
            for k in keep:
                if k in hdr: f.write(f"| {k} | {r[hdr.index(k)]} | {units[hdr.index(k)]} |\n")
            f.write("\n")
