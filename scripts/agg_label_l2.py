"""Merge the DRAM bytes of an `ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum -k regex:spmm` pass over
`bench_agg.py --quick --one-launch` (one launch per point) with the points' algorithmic bytes:
    l2_served = measured DRAM bytes < 0.8 x algorithmic bytes   (the gathers were partly served by L2)
usage: python scripts/agg_label_l2.py gpurun_out/agg_dram_r02.csv gpurun_out/agg_one_launch.jsonl > profiles/r02_agg_dram_labels.jsonl"""
import csv, json, sys
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10]
hdr = rows[0]
ik, im, iv, iid, iu = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("ID"), hdr.index("Metric Unit")
mult = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1, "us": 1e3, "ms": 1e6}
launches = {}
for r in rows[1:]:
    d = launches.setdefault(int(r[iid]), {"kernel": r[ik]})
    d[r[im]] = float(r[iv].replace(",", "")) * mult.get(r[iu], 1)
seq = [launches[k] for k in sorted(launches)]
pts = [json.loads(l) for l in open(sys.argv[2]) if l.startswith("{")]
pos = 0
def take_call(split):
    """kernels of one ops.spmm_csr call: [find_long_rows] spmm_csr_kernel [chunk combine]"""
    global pos
    out = []
    if split:
        while pos < len(seq) and "spmm_csr_kernel" not in seq[pos]["kernel"]:
            pos += 1                                    # (find_long_rows is not matched by the -k regex; defensive)
    out.append(seq[pos]); pos += 1
    if split:
        while pos < len(seq) and ("spmm_chunk" in seq[pos]["kernel"] or "spmm_combine" in seq[pos]["kernel"]):
            out.append(seq[pos]); pos += 1
    return out
for p in pts:
    if "error" in p:
        continue
    compare = p["family"].startswith("indeg") or p["family"] == "qws_cousage"   # these points also ran an unsplit launch
    split = compare or p["E"] >= (1 << 22)                 # ops.spmm_csr splits by default from 4M edges
    ks = take_call(split)
    if compare:
        take_call(False)                                # the unsplit comparison launch
    dram = sum(k.get("dram__bytes_read.sum", 0) + k.get("dram__bytes_write.sum", 0) for k in ks)
    print(json.dumps({"family": p["family"], "E": p["E"], "N": p["N"], "F": p["F"], "algorithmic_bytes": p["algorithmic_bytes"],
                      "dram_bytes_ncu": dram, "dram_over_algorithmic": dram / p["algorithmic_bytes"],
                      "l2_served": bool(dram < 0.8 * p["algorithmic_bytes"]), "gather_matrix_MB": p["gather_matrix_MB"]}))
