"""DRAM bytes of the GEMM launches of ONE bench step from an `ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum
-k regex:gemm_tc_kernel --csv` log of `bench.py --steps 2 --warmup 1 --no-graph ...` -> the JSON bench.py reads as roofline.traffic.
usage: python scripts/summarize_traffic.py traffic.csv launches_per_step out.json"""
import collections, csv, json, sys

path, per_step, out = sys.argv[1], int(sys.argv[2]), sys.argv[3]
lines = [l for l in open(path) if not l.startswith("==")]
by_id = collections.OrderedDict()
for r in csv.DictReader(lines):
    d = by_id.setdefault(r["ID"], {})
    v = float(r["Metric Value"].replace(",", ""))
    unit = r["Metric Unit"].lower()
    mult = {"byte": 1.0, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}.get(unit, 1.0)
    d[r["Metric Name"]] = v * mult
rows = list(by_id.values())[-per_step:]
rd = sum(r.get("dram__bytes_read.sum", 0.0) for r in rows)
wr = sum(r.get("dram__bytes_write.sum", 0.0) for r in rows)
json.dump({"launches": len(rows), "dram_read_bytes": rd, "dram_write_bytes": wr, "dram_bytes_per_launch": (rd + wr) / max(1, len(rows)),
           "source": "ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum -k regex:gemm_tc_kernel over `bench.py --steps 2 --warmup 1 --no-graph` "
                     "(VidVRD 200 videos, tf32+bf16x2); last %d launches = one step" % len(rows)}, open(out, "w"), indent=1)
print(open(out).read())
