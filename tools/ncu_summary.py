"""Summarise an .ncu-rep (from `ncu --set full`) into a small CSV for profiles/.
usage: python tools/ncu_summary.py gpurun_out/prof.ncu-rep profiles/r01_xxx.csv"""
import csv
import subprocess
import sys

WANT = ['Kernel Name', 'launch__grid_size', 'launch__block_size', 'launch__registers_per_thread',
        'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'launch__waves_per_multiprocessor', 'launch__occupancy_limit_shared_mem', 'launch__occupancy_limit_registers',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'lts__t_sector_hit_rate.pct',
        'sm__inst_executed_pipe_uniform.sum']


def main(rep, out):
    raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    idx = [(w, hdr.index(w)) for w in WANT if w in hdr]
    with open(out, 'w', newline='') as f:
        w = csv.writer(f)
        w.writerow([f'{n} [{units[i]}]' if units[i] else n for n, i in idx])
        for r in rows[2:]:
            w.writerow([r[i][:90] for _, i in idx])
    print(f'{out}: {len(rows) - 2} kernels')


if __name__ == '__main__':
    main(sys.argv[1], sys.argv[2])
