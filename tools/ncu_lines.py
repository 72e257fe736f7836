"""Dev tool: hottest CUDA source lines of ONE kernel of an .ncu-rep (ncu --set full --import-source on; kernels built with
-lineinfo): warp-stall samples, warp instructions executed, SASS instructions generated and the top stall reasons per line.
    python tools/ncu_lines.py gpurun_out/x.ncu-rep kernel-name-regex [top] > profiles/x_lines.md"""
import csv
import io
import re
import subprocess
import sys


def main():
    rep, want = sys.argv[1], re.compile(sys.argv[2])
    top_n = int(sys.argv[3]) if len(sys.argv) > 3 else 25
    raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                         capture_output=True, text=True).stdout
    blocks, cur = [], None
    for x in csv.reader(io.StringIO(raw)):
        if x and x[0] == "Function Name":
            cur = {"name": x[1], "rows": []}
            blocks.append(cur)
        elif cur is not None and x:
            cur["rows"].append(x)
    cand = [b for b in blocks if want.search(b["name"])]
    if not cand:
        sys.exit("no such kernel in the report")
    b = max(cand, key=lambda b: len(b["rows"]))
    h = b["rows"][0]
    si = h.index("# Samples")
    ei = h.index("Instructions Executed")
    stalls = [(i, n[6:]) for i, n in enumerate(h) if n.startswith("stall_") and "Not Issued" not in n]

    def num(v):
        try:
            return int(v)
        except ValueError:
            return 0

    lines, agg, cur_line, sass = {}, {}, None, 0
    for x in b["rows"][1:]:
        if x[0].isdigit():
            cur_line = int(x[0])
            st = {n: num(x[i]) for i, n in stalls if num(x[i])}
            lines[cur_line] = dict(src=x[1].strip(), samples=num(x[si]), ex=num(x[ei]), st=st, sass=0)
            for k, v in st.items():
                agg[k] = agg.get(k, 0) + v
        elif x[0] == "" and cur_line is not None:
            lines[cur_line]["sass"] += 1
            sass += 1
    tot = sum(l["samples"] for l in lines.values())
    name = b["name"].split("(")[0].split("::")[-1]
    print(f"# hottest source lines: `{name}` ({rep.split('/')[-1]})\n")
    print(f"{tot} warp-stall samples (SASS column: instructions attributed to the line, inlined callers included).  Stall reasons over the whole kernel: "
          + ", ".join(f"{k} {100 * v / max(1, sum(agg.values())):.0f} %" for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:7]) + ".\n")
    print("| line | share of samples | warp instructions | SASS | top stall reasons | source |")
    print("|---|---|---|---|---|---|")
    for ln, l in sorted(lines.items(), key=lambda kv: -kv[1]["samples"])[:top_n]:
        top = " ".join(f"{k}={v}" for k, v in sorted(l["st"].items(), key=lambda kv: -kv[1])[:3])
        src = l["src"][:110].replace("|", "\\|")
        print(f"| {ln} | {100 * l['samples'] / max(1, tot):.1f} % | {l['ex']} | {l['sass']} | {top} | `{src}` |")


if __name__ == "__main__":
    main()
