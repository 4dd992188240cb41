#!/usr/bin/env python
"""Summarise an `ncu --page source --csv` export: top SASS lines by stall samples + stall mix."""
import csv, sys
def main(path, top=25):
    rows = list(csv.reader(open(path)))
    hdr = rows[1]
    ix = {h: i for i, h in enumerate(hdr)}
    s_all = ix["Warp Stall Sampling (All Samples)"]
    stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    data = []
    tot = 0
    mix = {h: 0 for h in stalls}
    for r in rows[2:]:
        if len(r) < len(hdr): continue
        if r[s_all] == hdr[s_all]: break   # a second captured launch starts here: report the first only
        n = int(float(r[s_all] or 0))
        tot += n
        for h in stalls: mix[h] += int(float(r[ix[h]] or 0))
        data.append((n, r))
    print(rows[0][1][:100])
    print("stall mix:", ", ".join(f"{k[6:]}={100*v/max(tot,1):.0f}%" for k, v in sorted(mix.items(), key=lambda t: -t[1])[:8]))
    data.sort(key=lambda t: -t[0])
    for n, r in data[:top]:
        top_st = sorted(((int(float(r[ix[h]] or 0)), h[6:]) for h in stalls), reverse=True)[0]
        print(f"{100*n/max(tot,1):5.1f}%  {r[ix['Source']].strip()[:80]:80s} [{top_st[1]}]")
if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 25)
