#!/usr/bin/env python
"""Summarise an `ncu --page raw --csv` dump into the per-kernel-family JSON bench.py reads (profiles/rNN*_traffic.json).

    ncu -i gpurun_out/prof.ncu-rep --page raw --csv > profiles/r02_top_kernels_raw.csv
    python tools/ncu_summarize.py profiles/r02_top_kernels_raw.csv profiles/r02_traffic.json [--note "..."]

Per family: launches captured, mean ms, DRAM read / write / total bytes per launch, SM throughput, issue-active, warp occupancy,
tensor pipe, L1 LSU wavefronts, DRAM throughput (all % of peak), registers.  Families follow bench.py's profile tags.
"""
import csv
import json
import re
import sys

FAMILIES = [
    (r"attn_(tc|h16)_bwd_kernel", "attn_tc_bwd"), (r"attn_(tc|h16)_fwd_kernel", "attn_tc_fwd"), (r"attn_mma_bwd", "attn_umma_bwd"),
    (r"attn_mma_fwd", "attn_umma_fwd"), (r"ffn_fwd_kernel", "ffn_fwd"), (r"ffn_bwd1_kernel", "ffn_bwd1"), (r"ffn_bwd2_kernel", "ffn_bwd2"),
    (r"ffn_bwd_kernel", "ffn_bwd"), (r"lin_fwd_kernel<192", "inproj_fwd"), (r"lin_fwd_kernel<64", "outproj_norm_fwd"),
    (r"lin_bwd_kernel<192", "inproj_bwd"), (r"lin_bwd_kernel<64", "outproj_norm_bwd"), (r"fm_fwd_kernel", "fm_fused_fwd"),
    (r"fm_bwd_kernel", "fm_fused_bwd"), (r"gemm_tma_kernel", "gemm_tma"), (r"gemm_tc_ws_kernel", "gemm_tc_ws"), (r"gemm_tc_kernel", "gemm_tc"),
    (r"retrieve_fwd", "retrieve_fwd"), (r"retrieve_bwd", "retrieve_bwd"), (r"rowopt_catchup", "rowopt_catchup"),
    (r"rowopt_update", "rowopt_update"), (r"cat_bwd", "cat_bwd"), (r"cat_fwd", "cat_fwd"), (r"rowrms_bwd", "rowrms_bwd"),
    (r"rowrms_fwd", "rowrms_fwd"), (r"scatter_reduce", "scatter_reduce"),
]
COLS = {"ms": "gpu__time_duration.sum", "rd": "dram__bytes_read.sum", "wr": "dram__bytes_write.sum",
        "sm": "sm__throughput.avg.pct_of_peak_sustained_elapsed", "issue": "sm__issue_active.avg.pct_of_peak_sustained_elapsed",
        "warps": "sm__warps_active.avg.pct_of_peak_sustained_active", "tensor": "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "lsu": "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "dram": "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "regs": "launch__registers_per_thread"}
SCALE = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ms": 1.0, "us": 1e-3, "s": 1e3, "ns": 1e-6}


def main():
    raw, out = sys.argv[1], sys.argv[2]
    note = sys.argv[4] if len(sys.argv) > 4 and sys.argv[3] == "--note" else None
    rows = list(csv.reader(open(raw, newline="")))
    hi = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
    hdr, units, data = rows[hi], rows[hi + 1], rows[hi + 2:]
    ix = {k: next((i for i, h in enumerate(hdr) if h == v or h.endswith("." + v)), None) for k, v in COLS.items()}
    kn = hdr.index("Kernel Name")
    fam = {}
    for r in data:
        if len(r) <= kn:
            continue
        name = r[kn]
        f = next((t for pat, t in FAMILIES if re.search(pat, name)), None)
        if f is None:
            f = re.sub(r"\(.*", "", name).replace("void ", "").replace("b200ctr::", "").strip()

        def val(k):
            i = ix[k]
            if i is None or r[i] == "":
                return None
            try:
                return float(r[i].replace(",", "")) * SCALE.get(units[i], 1.0)
            except ValueError:                                   # "no data" for metrics a pass did not collect
                return None
        fam.setdefault(f, []).append({k: val(k) for k in COLS})
    res = {}
    for f, ls in fam.items():
        def mean(k):
            v = [x[k] for x in ls if x[k] is not None]
            return sum(v) / len(v) if v else None
        rd, wr = mean("rd") or 0.0, mean("wr") or 0.0
        res[f] = {"launches_captured": len(ls), "ncu_ms": round(mean("ms") or 0.0, 4), "dram_read_MB": round(rd / 1e6, 1), "dram_write_MB": round(wr / 1e6, 1),
                  "traffic_bytes_per_launch": int(rd + wr), "sm_throughput_pct": mean("sm"), "issue_active_pct": mean("issue"),
                  "warps_active_pct": mean("warps"), "tensor_pipe_pct": mean("tensor"), "l1_lsu_wavefronts_pct": mean("lsu"),
                  "dram_throughput_pct": mean("dram"), "registers": mean("regs")}
    if note:
        res["_note"] = note
    json.dump(res, open(out, "w"), indent=1)
    for f, v in sorted(res.items(), key=lambda kv: -(kv[1]["ncu_ms"] if isinstance(kv[1], dict) else 0)):
        if isinstance(v, dict):
            print(f"{f:22s} n={v['launches_captured']:3d} {v['ncu_ms']:8.4f} ms  dram {v['traffic_bytes_per_launch'] / 1e6:8.1f} MB  issue {v['issue_active_pct'] or 0:5.1f}%  "
                  f"tensor {v['tensor_pipe_pct'] or 0:5.1f}%  dram {v['dram_throughput_pct'] or 0:5.1f}%  regs {v['registers']}")


if __name__ == "__main__":
    main()
