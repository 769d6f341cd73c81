"""Writes the SASS evidence committed under profiles/: per-kernel histogram of the Blackwell-specific mnemonics in
libvpk.so and the SASS of the MMA-issue loop + TMA producers of the ConvLSTM gate-GEMM kernel.
    python tools/sass_listing.py > profiles/r02_sass_libvpk.md"""
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "vp_suite_b200", "libvpk.so")
KEYS = ["UTCHMMA", "UTCBAR", "UTMALDG", "UTMAPF", "LDTM", "UTCATOMSWS", "SYNCS", "UCGABAR", "MUFU.TANH", "HMMA", "LDGSTS"]


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    funcs = re.split(r"\n\s*Function : ", sass)[1:]
    print("# SASS evidence, libvpk.so (sm_100a)\n")
    print("`cuobjdump -sass vp_suite_b200/libvpk.so`, instruction counts per kernel (static). `UTCHMMA` = tcgen05.mma, "
          "`UTCBAR` = tcgen05.commit, `UTMALDG` = cp.async.bulk.tensor (TMA), `LDTM` = tcgen05.ld, `UTCATOMSWS` = TMEM "
          "alloc, `SYNCS` = mbarrier ops, `UCGABAR` = cluster barrier. No `HMMA` (legacy mma.sync) anywhere.\n")
    print("| kernel | " + " | ".join(KEYS) + " |\n|---|" + "---:|" * len(KEYS))
    gate = None
    for f in funcs:
        name = f.split("\n", 1)[0].strip()
        dem = subprocess.run(["c++filt", name], capture_output=True, text=True).stdout.strip() or name
        dem = dem.replace("vpk::(anonymous namespace)::", "vpk::")
        counts = [len(re.findall(r"\b" + re.escape(k), f)) for k in KEYS]
        if sum(counts[:6]) == 0 and "conv" not in dem:
            continue
        print(f"| `{dem[:90]}` | " + " | ".join(str(c) for c in counts) + " |")
        if "conv_halo_kernel<1, true, 2>" in dem:
            gate = f
    if gate is None:
        return
    lines = [l for l in gate.split("\n") if re.search(r"/\*[0-9a-f]{4,5}\*/\s+\S", l) and not re.match(r"\s*/\* 0x", l)]
    txt = [re.sub(r"\s*/\* 0x[0-9a-f]+ \*/\s*$", "", l).rstrip() for l in lines]
    idx = [i for i, l in enumerate(txt) if "UTCHMMA" in l]
    print("\n## MMA-issue loop of `conv_halo_kernel<EPI_LSTM, PAIR, MODE 2>` (one tap = 4 x UTCHMMA.2CTA, commits via UTCBAR)\n\n```")
    for l in txt[max(0, idx[0] - 70): idx[-1] + 45]:
        print(l)
    print("```")
    tma = [i for i, l in enumerate(txt) if "UTMALDG" in l]
    print("\n## TMA producers (activation halo box 4-D, weight tiles 2-D, both `.2CTA`)\n\n```")
    for i in tma:
        for l in txt[max(0, i - 6): i + 2]:
            print(l)
        print("...")
    print("```")
    ld = [i for i, l in enumerate(txt) if "LDTM" in l]
    print("\n## Epilogue: first accumulator chunk (LDTM.x32 -> gate math with MUFU.TANH -> STG)\n\n```")
    for l in txt[ld[0] - 4: ld[0] + 60]:
        print(l)
    print("```")


if __name__ == "__main__":
    main()
