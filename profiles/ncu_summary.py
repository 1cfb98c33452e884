"""Headline metrics of one kernel from `ncu -i <rep> --page raw --csv`.
usage: ncu_summary.py <raw.csv> <units per launch, e.g. particles*steps> [steps per launch]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
units_per_launch = float(sys.argv[2]); steps = float(sys.argv[3]) if len(sys.argv) > 3 else 1.0
hdr, units, vals = rows[0], rows[1], rows[2]
want = ['Kernel Name', 'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread', 'launch__grid_size',
        'launch__block_size', 'lts__t_sector_hit_rate.pct', 'l1tex__t_sector_hit_rate.pct',
        'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
        'smsp__sass_inst_executed_op_local_ld.sum', 'smsp__sass_inst_executed_op_local_st.sum']
d = {}
for w in want:
    if w in hdr:
        i = hdr.index(w)
        print(f"{w:88s} {vals[i]} {units[i]}")
        d[w] = (vals[i], units[i])


def to_bytes(v, u):
    return float(v) * {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}.get(u, 1.0)


def to_us(v, u):
    return float(v) * {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(u, 1.0)


traffic = to_bytes(*d['dram__bytes_read.sum']) + to_bytes(*d['dram__bytes_write.sum'])
inst = float(d['smsp__inst_executed.sum'][0])
print(f"# per unit: DRAM {traffic / units_per_launch:.1f} B, {inst * 32 / units_per_launch:.0f} thread instructions; "
      f"{to_us(*d['gpu__time_duration.sum']) / steps:.0f} us per step under the profiler")
