"""Top SASS instructions by stall samples from `ncu --page source --csv` output (one or more kernels)."""
import csv, sys, collections
path, pat = sys.argv[1], (sys.argv[2] if len(sys.argv) > 2 else "")
topn = int(sys.argv[3]) if len(sys.argv) > 3 else 40
rows = list(csv.reader(open(path)))
i = 0
while i < len(rows):
    r = rows[i]
    if r and r[0] == "Kernel Name":
        name = r[1]; hdr = rows[i + 1]; j = i + 2
        body = []
        while j < len(rows) and not (rows[j] and rows[j][0] == "Kernel Name"):
            if rows[j]: body.append(rows[j])
            j += 1
        if pat in name:
            H = {h: k for k, h in enumerate(hdr)}
            tot = sum(int(b[H["# Samples"]] or 0) for b in body)
            inst = sum(int(b[H["Instructions Executed"]] or 0) for b in body)
            print(f"=== {name}\n    samples={tot} warp-inst={inst} sass-lines={len(body)}")
            reasons = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
            agg = {h: sum(int(b[H[h]] or 0) for b in body) for h in reasons}
            print("    " + " ".join(f"{k[6:]}={v}" for k, v in sorted(agg.items(), key=lambda kv: -kv[1]) if v))
            opc = collections.Counter()
            for b in body:
                op = b[H["Source"]].split()[0] if b[H["Source"]].split() else "?"
                if op.startswith("@"): op = b[H["Source"]].split()[1]
                opc[op.split(".")[0]] += int(b[H["Instructions Executed"]] or 0)
            print("    opcodes: " + " ".join(f"{k}={v*100//max(inst,1)}%" for k, v in opc.most_common(14)))
            idx = sorted(range(len(body)), key=lambda k: -int(body[k][H["# Samples"]] or 0))[:topn]
            for k in sorted(idx):
                b = body[k]
                st = {h[6:]: int(b[H[h]] or 0) for h in reasons}
                top = sorted(st.items(), key=lambda kv: -kv[1])[:2]
                print(f"    [{k:5d}] smp={int(b[H['# Samples']] or 0):6d} exec={int(b[H['Instructions Executed']] or 0):9d} {b[H['Source']].strip()[:70]:70s} {top}")
        i = j
    else:
        i += 1
