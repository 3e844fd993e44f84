"""Aggregate an ncu `--metrics gpu__time_duration.sum --csv` launch list by kernel: count, total us, share."""
import collections, csv, re, sys


def main(path, top=40):
    rows = list(csv.reader(open(path, errors="ignore")))
    hdr, data = None, []
    for r in rows:
        if hdr is None:
            if "Kernel Name" in r:
                hdr = r
            continue
        data.append(dict(zip(hdr, r)))
    agg = collections.defaultdict(lambda: [0, 0.0])
    for d in data:
        if d.get("Metric Name") != "gpu__time_duration.sum":
            continue
        name = re.sub(r"\(.*", "", d["Kernel Name"]).replace("void ", "").replace("cdae::", "")
        t = float(d["Metric Value"].replace(",", ""))
        t = t / 1000 if d["Metric Unit"] == "ns" else t * 1000 if d["Metric Unit"] == "ms" else t
        agg[name][0] += 1; agg[name][1] += t
    tot = sum(v[1] for v in agg.values())
    print(f"{sum(v[0] for v in agg.values())} launches, {tot:.1f} us (serialised, cold-cache ncu times)")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
        print(f"{v[1]:10.1f} us {v[0]:5d}  {100 * v[1] / tot:5.1f}%  {k[:100]}")


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 40)
