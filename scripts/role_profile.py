"""Per-role view of an `ncu --set full --import-source on` capture of md2_march_roles: instructions per image row, share
of the warp-state samples and stall mix of each role's loop (the loops are delimited by their BAR.SYNC), plus the most
stalled instructions.   usage: role_profile.py prof.ncu-rep [top_n]"""
import csv, io, subprocess, sys

rep = sys.argv[1]
topn = int(sys.argv[2]) if len(sys.argv) > 2 else 0
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
R = [r for r in rows[2:] if len(r) >= len(hdr)]
f = lambda r, k: float(r[ix[k]] or 0)
ex = [f(r, "Instructions Executed") for r in R]
rows_per_launch = max(e for i, e in enumerate(ex) if "BAR.SYNC" in R[i][ix["Source"]])
hot = [i for i, e in enumerate(ex) if e >= 0.4 * rows_per_launch]
bars = [i for i in hot if "BAR.SYNC" in R[i][ix["Source"]]]
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
tot = sum(f(r, "# Samples") for r in R)
print("kernel:", rows[0][1][:110])
print("loop trips (executed count of the row barrier): %d; total samples %d" % (rows_per_launch, tot))
start = hot[0]
for n, b in enumerate(bars):
    end = b
    while end + 1 < len(R) and ex[end + 1] >= 0.4 * rows_per_launch and "BAR.SYNC" not in R[end + 1][ix["Source"]] and end - b < 6:
        end += 1          # the loop-closing branch after the barrier carries the barrier-wait samples
    rr = R[start:end + 1]
    inst = sum(f(r, "Instructions Executed") for r in rr)
    smp = sum(f(r, "# Samples") for r in rr)
    st = sorted(((s, sum(f(r, s) for r in rr)) for s in stalls), key=lambda kv: -kv[1])[:7]
    # role A's loop holds two rows per trip since MD2_ROLE_A_AHEAD=2: four barrier-delimited bodies = A even, A odd, B, C
    # (the per-row figures of the two A halves are per TWO rows of the launch: add them for A's instructions per row)
    names = ["A (even rows)", "A (odd rows)", "B", "C"] if len(bars) == 4 else ["A", "B", "C"]
    print("role %s: %d static, %.0f warp-instr per row, %.1f %% of samples: %s" % (
        names[n] if n < len(names) else str(n), len(rr), inst / rows_per_launch, 100 * smp / tot,
        " ".join("%s=%.0f%%" % (k[6:], 100 * v / max(smp, 1)) for k, v in st)))
    if topn:
        for i in sorted(sorted(range(start, end + 1), key=lambda i: -f(R[i], "# Samples"))[:topn]):
            r = R[i]
            print("    %5d %-64s smp %5s long %5s short %5s wait %4s bar %5s" % (
                i, r[ix["Source"]][:64], r[ix["# Samples"]], r[ix["stall_long_sb"]], r[ix["stall_short_sb"]],
                r[ix["stall_wait"]], r[ix["stall_barrier"]]))
    start = end + 1
