import csv, subprocess, sys, io, collections
rep = sys.argv[1]
raw = subprocess.run(['ncu','-i',rep,'--page','raw','--csv'],capture_output=True,text=True).stdout
r = list(csv.reader(io.StringIO(raw)))
hdr = r[0]
want = ['Kernel Name','gpu__time_duration.sum','launch__grid_size','launch__registers_per_thread','launch__occupancy_limit_shared_mem','sm__warps_active.avg.pct_of_peak_sustained_active','dram__bytes_read.sum','dram__bytes_write.sum','gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed','sm__throughput.avg.pct_of_peak_sustained_elapsed','sm__inst_executed_pipe_tensor.sum','sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active','sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active','smsp__cycles_active.avg','sm__cycles_elapsed.avg','lts__t_bytes.sum','l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum','smsp__inst_executed.sum','sm__clocks_per_second' if False else 'sm__cycles_elapsed.avg.per_second']
for w in want:
    if w in hdr:
        i = hdr.index(w); print(f'{w:75s}', [row[i] for row in r[1:]][:4])
tens = [h for h in hdr if 'tensor' in h]
print('tensor metrics:', tens[:12])
# source page: top stall lines
src = subprocess.run(['ncu','-i',rep,'--page','source','--csv','--print-source','cuda'] ,capture_output=True,text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
if rows:
    h = rows[0]
    def col(name):
        for i,x in enumerate(h):
            if x.strip()==name: return i
        return None
    ci = col('Source'); cs = col('# Samples') or col('Samples'); 
    print([x for x in h][:30])
