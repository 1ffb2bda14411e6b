import csv,sys,subprocess
rep=sys.argv[1]
out=subprocess.run(["ncu","-i",rep,"--page","raw","--csv"],capture_output=True,text=True).stdout
rows=list(csv.reader(out.splitlines()))
hdr=rows[0]
if len(sys.argv)>2 and sys.argv[2]=="names":
    for h in hdr:
        if any(k in h for k in sys.argv[3:]): print(h)
    sys.exit()
cols=["Kernel Name","Grid Size","gpu__time_duration.sum","sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active","sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed","dram__bytes_read.sum","dram__bytes_write.sum","dram__throughput.avg.pct_of_peak_sustained_elapsed","lts__throughput.avg.pct_of_peak_sustained_elapsed","l1tex__throughput.avg.pct_of_peak_sustained_active","sm__warps_active.avg.pct_of_peak_sustained_active","launch__registers_per_thread","launch__occupancy_limit_shared_mem","launch__occupancy_limit_registers"]
idx=[hdr.index(c) if c in hdr else -1 for c in cols]
print(" | ".join(cols))
print("units:", " | ".join(rows[1][i] if i>=0 else "-" for i in idx))
for r in rows[2:]:
    print(" | ".join((r[i][:44] if i>=0 else "-") for i in idx))
