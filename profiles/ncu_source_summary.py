"""Summarise an `ncu --set full --import-source on` report by CUDA source line: stall samples, executed
instructions and the dominant stall reasons.   python profiles/ncu_source_summary.py report.ncu-rep [top]"""
import csv, subprocess, sys, io


def I(x):
    try:
        return int(x)
    except ValueError:
        return 0


def main():
    rep = sys.argv[1]
    top = int(sys.argv[2]) if len(sys.argv) > 2 else 14
    raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    cur, files, H, hdr = None, {}, None, {}
    for r in rows:
        if r and r[0] == "File Path":
            cur = r[1]; files[cur] = []; continue
        if r and r[0] == "Line No":
            H = r
            hdr[cur] = r
            continue
        if r and r[0] == "Function Name":
            continue
        if cur and len(r) > 8:
            files[cur].append(r)
    tot = {}
    allsamp = allinst = 0
    for f, rs in files.items():
        Hf = hdr.get(f, H)                                   # every file section has its own header row
        stf = [i for i, h in enumerate(Hf) if h.startswith("stall_") and "Not Issued" not in h]
        for r in rs:
            if r[0].strip().isdigit() and len(r) == len(Hf):
                allsamp += I(r[4]); allinst += I(r[7])
                for i in stf:
                    tot[Hf[i][6:]] = tot.get(Hf[i][6:], 0) + I(r[i])
    st = [i for i, h in enumerate(H) if h.startswith("stall_") and "Not Issued" not in h]
    print(f"samples {allsamp}  warp instructions {allinst}")
    print("stall reasons:", {k: v for k, v in sorted(tot.items(), key=lambda kv: -kv[1]) if v})
    for f, rs in files.items():
        src = [r for r in rs if r[0].strip().isdigit() and I(r[4]) > 0]
        if not src or sum(I(r[4]) for r in src) < allsamp * 0.02:
            continue
        print("==", f.split("/")[-1], "samples", sum(I(r[4]) for r in src), "instr", sum(I(r[7]) for r in src))
        for r in sorted(src, key=lambda r: -I(r[4]))[:top]:
            why = {H[i][6:]: I(r[i]) for i in st if I(r[i]) * 8 > I(r[4])}
            print(f"  {r[0]:>4s} {I(r[4]):6d} {I(r[7]):9d}  {r[1].strip()[:80]:80s} {why}")


if __name__ == "__main__":
    main()
