#!/usr/bin/env python
"""Source-page aggregation of a K4 (voice_render_mix_tma) capture: who waits for whom.
usage: tools/ncu_roles.py gpurun_out/prof.ncu-rep >> profiles/<summary>.txt

The kernel is one producer warp + eight consumer warps.  The SASS of the producer precedes the consumer loop; the slow
paths of the mbarrier waits (NANOSLEEP loops) are laid out behind the kernel body.  The consumer loop is recognised by its
full-barrier wait (a TRYWAIT on [R+URZ] without an offset), the producer's waits by the empty barriers' offset."""
import csv
import re
import subprocess
import sys


def main():
    rep = sys.argv[1]
    txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    hdr, data = rows[1], rows[2:]
    ix = {h: i for i, h in enumerate(hdr)}

    def f(r, k):
        try:
            return float(r[ix[k]])
        except (ValueError, IndexError):
            return 0.0

    src = [r[ix["Source"]].strip() for r in data]
    full_waits = [i for i, s in enumerate(src) if re.search(r"TRYWAIT P\d, \[R\d+\+URZ\],", s)]
    sleeps = [i for i, s in enumerate(src) if "NANOSLEEP" in s]
    if not full_waits or not sleeps:
        print("# (no warp-specialised layout recognised)")
        return
    c0, c_ool = full_waits[0], full_waits[-1]
    body_end = sleeps[0] - 1                      # first out-of-line wait loop
    tot_s = sum(f(r, "# Samples") for r in data)
    tot_i = sum(f(r, "Instructions Executed") for r in data)

    def region(idx):
        return sum(f(data[i], "# Samples") for i in idx), sum(f(data[i], "Instructions Executed") for i in idx)

    prod = list(range(0, c0 - 30))
    cons = list(range(c0 - 30, body_end))
    prod_wait = [i for i in range(body_end, c_ool - 1)]
    cons_wait_ool = list(range(c_ool - 1, len(data)))
    bra = next((i for i in range(c0, c0 + 8) if re.search(r"^@!P\d\s+BRA", src[i])), c0 + 1)   # the wait's branch collects its samples
    cons_wait_inline = list(range(c0, bra + 1))
    ps, pi = region(prod)
    cs, ci = region(cons)
    pws, pwi = region(prod_wait)
    cws, cwi = region(cons_wait_ool)
    cis, _ = region(cons_wait_inline)
    print(f"\n# source page (ncu --page source), aggregated over the SASS of the kernel: {int(tot_s)} stall samples, "
          f"{tot_i / 1e6:.0f} M executed warp instructions")
    print(f"#   producer warp: {ps / tot_s * 100:.1f} % of the samples at work ({pi / 1e6:.0f} M instructions) + "
          f"{pws / tot_s * 100:.1f} % waiting for an empty stage")
    print(f"#   consumer warps: {(cs - cis) / tot_s * 100:.1f} % of the samples at work ({ci / 1e6:.0f} M instructions) + "
          f"{(cis + cws) / tot_s * 100:.1f} % waiting for a full stage ({cwi / 1e6:.0f} M instructions of the wait loop)")
    share = (cis + cws) / max(cs + cws, 1.0)
    print(f"#   -> a consumer warp waits {share * 100:.1f} % of its time; the producer warp waits "
          f"{pws / max(ps + pws, 1.0) * 100:.1f} % of its time")


if __name__ == "__main__":
    main()
