"""Summarise .ncu-rep files (ncu --set full) into one CSV: python tools/ncu_summary.py a.ncu-rep b.ncu-rep > profiles/x.csv"""
import csv,sys,subprocess
want=['Kernel Name','launch__grid_size','launch__block_size','launch__registers_per_thread','launch__occupancy_limit_registers','gpu__time_duration.sum','dram__bytes_read.sum','dram__bytes_write.sum','gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed','lts__t_sector_hit_rate.pct','l1tex__t_sector_hit_rate.pct','sm__warps_active.avg.pct_of_peak_sustained_active','sm__inst_issued.avg.pct_of_peak_sustained_active','sm__throughput.avg.pct_of_peak_sustained_elapsed','smsp__cycles_active.avg','sm__cycles_elapsed.avg.per_second']
out=csv.writer(sys.stdout)
first=True
for rep in sys.argv[1:]:
    raw=subprocess.run(['ncu','-i',rep,'--page','raw','--csv'],capture_output=True,text=True).stdout
    rows=list(csv.reader(raw.splitlines()))
    hdr=rows[0]; units=rows[1]
    idx=[hdr.index(w) for w in want if w in hdr]
    if first:
        out.writerow(['report']+[hdr[i] for i in idx]); first=False
    for r in rows[2:]:
        out.writerow([rep.split('/')[-1]]+[(r[i][:90] + (' ' + units[i] if units[i] else '')) for i in idx])
