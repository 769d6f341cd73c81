"""Index table + readings in front of tools/ncu_summary.py's per-launch tables for the 17-launch cfg-5 step capture.
    python tools/ncu_summary.py gpurun_out/x.ncu-rep x > /tmp/full.md
    python tools/ncu_step_table.py /tmp/full.md > profiles/r02_ncu_cfg5_step_b64.md"""
import re
import sys

NAMES = ["encoder.rnn1 (gate GEMM, K = 9*80, N = 4 x 64)", "encoder.stage2.conv k3s2", "encoder.rnn2", "encoder.stage3.conv k3s2",
         "encoder.rnn3", "forecaster.rnn3 (h only)", "forecaster.stage3.deconv k4s2, parity 0", "... parity 1", "... parity 2",
         "... parity 3", "forecaster.rnn2", "forecaster.stage2.deconv k4s2, parity 0", "... parity 1", "... parity 2",
         "... parity 3", "forecaster.rnn1 (K = 9*160, N = 256: the dominant kernel)",
         "final_fused (stage1 deconv k3 + 1x1 head, MODE 3)"]


def g(sec, k):
    m = re.search(re.escape(k) + r"[^|]*\| ([0-9.]+) (\S*)", sec)
    return float(m.group(1)) if m else 0.0


def main():
    s = open(sys.argv[1]).read()
    secs = s.split("## launch ")[1:]
    rows = []
    for n, sec in zip(NAMES, secs):
        rows.append((n, sec.split("\n")[0].split("`")[1].replace("void unnamed>::", ""), g(sec, "kernel duration"),
                     g(sec, "tensor pipe active"), g(sec, "DRAM traffic read + write"), g(sec, "L2 -> SM bytes"),
                     g(sec, "DRAM throughput (of peak)"), g(sec, "registers per thread")))
    tot = sum(r[2] for r in rows)
    print("# r02: `ncu --set full` over one timestep's tcgen05 launches of the cfg-5 rollout (64 sequences, final build)\n")
    print("Command (on the B200 box): `ncu --set full --clock-control none --import-source on -k regex:conv_halo_kernel "
          "--launch-skip 45 --launch-count 17 -o gpurun_out/r02_ncu_cfg5_step_b64_v3 python tools/run_once.py cfg5 64` "
          "(`tools/r02_evidence.sh`).")
    print("17 consecutive launches = one of every layer shape of the rollout (cold caches, serialised replays: compare "
          "shares and ratios, not absolutes).\n")
    print("| # | layer | kernel | us | share | tensor pipe active | DRAM MB | DRAM % of peak | L2->SM GB | regs |")
    print("|---|---|---|---|---|---|---|---|---|---|")
    for i, r in enumerate(rows):
        print(f"| {i} | {r[0]} | `{r[1]}` | {r[2]:.1f} | {100 * r[2] / tot:.1f} % | {r[3]:.1f} % | {r[4]:.1f} | {r[6]:.1f} % | "
              f"{r[5]:.2f} | {int(r[7])} |")
    print(f"\nSum {tot:.0f} us.\n")
    sys.stdout.write(open(sys.argv[2]).read() if len(sys.argv) > 2 else "")
    print("\nFull per-launch metric tables follow.\n")
    print(s[s.index("## launch 0"):])


if __name__ == "__main__":
    main()
