"""Turns an .ncu-rep (ncu --set full --import-source on) into the markdown summary committed under profiles/.
    python tools/ncu_summary.py gpurun_out/x.ncu-rep "title" > profiles/x.md
Reads the report with `ncu -i ... --page raw/source --csv` (works without a GPU)."""
import csv
import io
import subprocess
import sys

METRICS = [
    ("gpu__time_duration.sum", "kernel duration"),
    ("launch__grid_size", "grid (CTAs)"),
    ("launch__block_size", "threads per CTA"),
    ("launch__registers_per_thread", "registers per thread"),
    ("launch__shared_mem_per_block_dynamic", "dynamic shared memory per CTA"),
    ("sm__cycles_elapsed.avg", "SM cycles elapsed"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "tensor pipe active (of elapsed)"),
    ("sm__inst_executed_pipe_tensor_subpipe_hmma.sum", "tensor instructions (UTCHMMA)"),
    ("dram__bytes_read.sum", "DRAM bytes read"),
    ("dram__bytes_write.sum", "DRAM bytes written"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput (of peak)"),
    ("l1tex__m_xbar2l1tex_read_bytes.sum", "L2 -> SM bytes"),
    ("l1tex__m_xbar2l1tex_read_bytes.sum.per_second", "L2 -> SM rate"),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 throughput (of peak)"),
    ("lts__t_sector_hit_rate.pct", "L2 hit rate"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy"),
    ("smsp__inst_executed.sum", "warp instructions executed"),
]


def ncu_csv(rep, page, extra=()):
    out = subprocess.run(["ncu", "-i", rep, "--page", page, "--csv", *extra], capture_output=True, text=True).stdout
    return list(csv.reader(io.StringIO(out)))


def main():
    rep, title = sys.argv[1], sys.argv[2]
    rows = ncu_csv(rep, "raw")
    hdr, units, data = rows[0], rows[1], rows[2:]
    name_i = hdr.index("Kernel Name")
    print(f"# {title}\n")
    print(f"Source: `{rep.split('/')[-1]}` (`ncu --set full --clock-control none --import-source on`; values are per launch, "
          "cold caches, serialised replays: compare shares and ratios, not absolutes).\n")
    for k, r in enumerate(data):
        print(f"## launch {k}: `{r[name_i][:100]}`\n")
        print("| metric | value |\n|---|---|")
        for m, label in METRICS:
            if m in hdr:
                i = hdr.index(m)
                print(f"| {label} (`{m}`) | {r[i]} {units[i]} |")
        try:
            rd = float(r[hdr.index("dram__bytes_read.sum")])
            wr = float(r[hdr.index("dram__bytes_write.sum")])
            u = units[hdr.index("dram__bytes_read.sum")]
            print(f"| DRAM traffic read + write | {rd + wr:.1f} {u} |")
        except (ValueError, IndexError):
            pass
        print()
    src = ncu_csv(rep, "source")
    if len(src) > 3:
        h = src[1]
        d = src[2:]
        ia, isrc, iex = h.index("Warp Stall Sampling (All Samples)"), h.index("Source"), h.index("Instructions Executed")

        def iv(x):
            try:
                return int(x)
            except ValueError:
                return 0
        # the source page lists every launch of the report back to back: keep the first
        first = []
        seen = set()
        for r in d:
            if len(r) <= ia:
                continue
            if r[0] in seen:
                break
            seen.add(r[0])
            first.append(r)
        tot = sum(iv(r[ia]) for r in first) or 1
        print("## warp-stall sampling, top SASS instructions (launch 0)\n")
        print("| samples | share | executed | SASS |\n|---:|---:|---:|---|")
        for r in sorted(first, key=lambda r: -iv(r[ia]))[:25]:
            print(f"| {iv(r[ia])} | {100 * iv(r[ia]) / tot:.1f}% | {r[iex]} | `{r[isrc].strip()[:90]}` |")
        mn = {}
        for r in first:
            op = r[isrc].strip().split()
            if not op:
                continue
            o = op[1] if op[0].startswith("@") and len(op) > 1 else op[0]
            o = o.split(".")[0]
            if o in ("UTCHMMA", "UTMALDG", "LDTM", "UTCBAR", "UTMASTG", "UBLKCP", "SYNCS", "MUFU", "HMMA"):
                mn[o] = mn.get(o, 0) + iv(r[iex])
        print("\nExecuted instruction counts of the Blackwell-specific mnemonics: " +
              ", ".join(f"`{k}` {v}" for k, v in sorted(mn.items())) + ".")


if __name__ == "__main__":
    main()
