"""Turn the ncu csv files of tools/ncu_fp64_ops.sh (gpurun_out/ops_<key>.csv) into profiles/fp64_ops.json, stamped with
the hash of the source tree they were captured on (tools/stamp.py): bench.py uses the executed-instruction counts only
while the stamp matches."""
import csv
import glob
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))
import stamp

NODES = {"cfg2": 1_000_000, "euler_L8_lrot": 1_000_000, "cfg5_step": 1_000_000, "cfg3": 1_000_000, "cfg4": 300_000,
         "cfg2_rnlm": 1_000_000, "euler_L8_lrot_rnlm": 1_000_000, "eij": 2_000_000}


def parse(path):
    rows = [r for r in csv.reader(l for l in open(path) if l.startswith('"'))]
    hdr = rows[0]
    mi, vi = hdr.index("Metric Name"), hdr.index("Metric Value")
    return {r[mi]: float(r[vi].replace(",", "")) for r in rows[1:] if len(r) > vi}


if __name__ == "__main__":
    src = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "gpurun_out")
    out = {"_comment": "FP64 thread instructions EXECUTED per node-step (per Eij evaluation for 'eij'), measured with ncu "
                       "(smsp__sass_thread_inst_executed_op_{dfma,dmul,dadd}_pred_on.sum of one launch / nodes; tools/ncu_fp64_ops.sh). "
                       "bench.py's FP64 roofline counts min(executed, code generator's count) pipe slots per node-step; pipe_fp64_pct is "
                       "ncu's sm__pipe_fp64_cycles_active for the same launch (a cross-check of roofline.frac, measured under the profiler).",
           "_stamp": stamp.tree_hash()}
    for f in sorted(glob.glob(os.path.join(src, "ops_*.csv"))):
        key = os.path.basename(f)[4:-4]
        try:
            m = parse(f)
        except Exception as ex:   # noqa
            print("skip", f, ex)
            continue
        n = NODES.get(key, 1_000_000)
        out[key] = {"dfma": round(m["smsp__sass_thread_inst_executed_op_dfma_pred_on.sum"] / n),
                    "dmul": round(m["smsp__sass_thread_inst_executed_op_dmul_pred_on.sum"] / n),
                    "dadd": round(m["smsp__sass_thread_inst_executed_op_dadd_pred_on.sum"] / n), "nodes": n,
                    "pipe_fp64_pct": m.get("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active"),
                    "warp_inst_executed": m.get("smsp__inst_executed.sum"), "warp_inst_fp64": m.get("sm__inst_executed_pipe_fp64.sum"),
                    "ncu_duration_us": m.get("gpu__time_duration.sum")}
    traffic = {"_comment": "DRAM bytes per node-step: (dram__bytes_read.sum + dram__bytes_write.sum) of one launch / nodes of that launch (ncu, same captures as fp64_ops.json); bench.py scales it to its own launch size", "_stamp": out["_stamp"]}
    for f in sorted(glob.glob(os.path.join(src, "ops_*.csv"))):
        key = os.path.basename(f)[4:-4]
        try:
            m = parse(f)
            nn = NODES.get(key, 1_000_000)
            traffic[key] = (m["dram__bytes_read.sum"] + m["dram__bytes_write.sum"]) / nn          # bytes per node (per evaluation for eij)
            traffic[key + "_read_write_nodes"] = [int(m["dram__bytes_read.sum"]), int(m["dram__bytes_write.sum"]), nn]
        except Exception:
            pass
    json.dump(traffic, open(os.path.join(ROOT, "profiles", "traffic.json"), "w"), indent=1)
    json.dump(out, open(os.path.join(ROOT, "profiles", "fp64_ops.json"), "w"), indent=1)
    print("wrote profiles/fp64_ops.json", sorted(k for k in out if not k.startswith("_")), "stamp", out["_stamp"])
