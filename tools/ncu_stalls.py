"""Dev tool: warp-stall samples of ONE kernel of an .ncu-rep (ncu --set full --import-source on), summed per
barrier-delimited SASS segment — which phase of a multi-phase kernel the warps spend their time in, and on what.
    python tools/ncu_stalls.py gpurun_out/x.ncu-rep [kernel-name-regex] > profiles/x_stalls.md"""
import csv
import io
import subprocess
import sys

MARKS = ("UTCHMMA", "LDTM", "UTCBAR", "SYNCS", "MUFU.TANH", "ATOMG", "ATOMS", "RED", "STG", "LDG", "MUFU.EX2", "MUFU.RCP", "SHFL")


def main():
    rep = sys.argv[1]
    import re
    cmd = ["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"]
    raw = subprocess.run(cmd, capture_output=True, text=True).stdout
    allrows = list(csv.reader(io.StringIO(raw)))
    # one block per profiled launch: ["Kernel Name", name], header, instruction rows
    starts = [i for i, x in enumerate(allrows) if x and x[0] == "Kernel Name"] + [len(allrows)]
    want = re.compile(sys.argv[2]) if len(sys.argv) > 2 else None
    r = None
    for a, b in zip(starts[:-1], starts[1:]):
        if want is None or want.search(allrows[a][1]):
            r = [x for x in allrows[a:b] if x]
            break
    if r is None:
        sys.exit("no such kernel in the report")
    name, h, rows = r[0][1].split("(")[0], r[1], r[2:]
    ix = {n: i for i, n in enumerate(h)}
    stalls = [n for n in h if n.startswith("stall_") and "Not Issued" not in n]
    tot = sum(int(x[ix["# Samples"]]) for x in rows)
    print(f"# warp-stall samples per barrier-delimited segment: `{name.split('::')[-1]}` ({rep.split('/')[-1]})\n")
    print(f"{tot} samples, {len(rows)} SASS instructions.  A segment ends at a BAR.SYNC; the wait AT a barrier is attributed to the "
          "segment that follows it (`barrier` stalls at the top of a segment = waiting for the slowest warp of the previous one).\n")
    print("| first SASS index | instructions | warp-instructions executed | samples | share | top stall reasons | what is in it |")
    print("|---|---|---|---|---|---|---|")
    cur = dict(start=0, samples=0, st={}, inst=0, marks=[], ex=0)
    segs = []
    for k, x in enumerate(rows):
        src = x[ix["Source"]].strip()
        cur["samples"] += int(x[ix["# Samples"]])
        cur["inst"] += 1
        cur["ex"] += int(x[ix["Instructions Executed"]])
        for n in stalls:
            v = int(x[ix[n]])
            if v:
                cur["st"][n] = cur["st"].get(n, 0) + v
        op = src.split()[1] if src.startswith("@") else src.split()[0]
        for m in MARKS:
            if op.startswith(m):
                if cur["marks"] and cur["marks"][-1][0] == m:
                    cur["marks"][-1][1] += 1
                else:
                    cur["marks"].append([m, 1])
        if op.startswith("BAR") or op.startswith("EXIT"):
            segs.append(cur)
            cur = dict(start=k + 1, samples=0, st={}, inst=0, marks=[], ex=0)
    segs.append(cur)
    for s in segs:
        if s["samples"] < tot * 0.003:
            continue
        top = sorted(s["st"].items(), key=lambda kv: -kv[1])[:4]
        print(f"| {s['start']} | {s['inst']} | {s['ex']} | {s['samples']} | {100 * s['samples'] / tot:.1f} % | "
              + " ".join(f"{k[6:]}={v}" for k, v in top) + " | " + " ".join(f"{m}×{c}" for m, c in s["marks"][:12]) + " |")


if __name__ == "__main__":
    main()
