#!/usr/bin/env python3
"""Summarise `ncu --set full` captures into profiles/<tag>_traffic.json (what bench.py's roofline block reads) and one
details / stalls text per kernel.

usage: python tools/ncu_summary.py <tag> name=path.ncu-rep [name=path.ncu-rep ...]
       e.g. python tools/ncu_summary.py r2u cs_wedge_kernel=gpurun_out/r2u_wedge_cfg2.ncu-rep cs_search2_kernel=...

The JSON carries `kernel_src_sha`: the hash of slam.net_b200/csrc at the time of the capture; bench.py refuses the numbers
(traffic = null) when the kernel sources have changed since."""
import csv, hashlib, io, json, os, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def kernel_src_sha():
    h = hashlib.sha256()
    d = os.path.join(ROOT, "slam.net_b200", "csrc")
    for f in sorted(os.listdir(d)):
        if f.endswith((".cu", ".cuh", ".h")):
            h.update(f.encode())
            h.update(open(os.path.join(d, f), "rb").read())
    return h.hexdigest()[:16]


def raw_metrics(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr = next(r for r in rows if r and r[0] == "ID")
    i = rows.index(hdr)
    units = rows[i + 1]
    vals = rows[i + 2] if len(rows) > i + 2 else []
    scale = {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "us": 1e3, "usecond": 1e3, "ms": 1e6, "msecond": 1e6, "s": 1e9, "second": 1e9}  # -> bytes, nanoseconds
    out_ = {}
    for k, u, v in zip(hdr, units, vals):
        if u in scale:
            try:
                v = repr(float(v.replace(",", "")) * scale[u])
            except ValueError:
                pass
        out_[k] = v
    return out_


def num(m, key, default=None):
    try:
        return float(m[key].replace(",", ""))
    except Exception:
        return default


if __name__ == "__main__":
    tag = sys.argv[1]
    res = {"source": "ncu --set full --clock-control none --import-source on; " + " ".join(sys.argv[2:]),
           "kernel_src_sha": kernel_src_sha(), "git_head": subprocess.run(["git", "rev-parse", "--short", "HEAD"], cwd=ROOT,
                                                                          capture_output=True, text=True).stdout.strip()}
    for spec in sys.argv[2:]:
        name, rep = spec.split("=", 1)
        m = raw_metrics(rep)
        res[name] = {
            "dram_bytes_read": num(m, "dram__bytes_read.sum"), "dram_bytes_write": num(m, "dram__bytes_write.sum"),
            "duration_us": (num(m, "gpu__time_duration.sum", 0.0) or 0.0) / 1e3,
            "warp_instructions": num(m, "smsp__inst_executed.sum"),
            "ipc_active": num(m, "sm__inst_executed.avg.per_cycle_active"),
            "issue_slots_busy_pct": num(m, "sm__inst_issued.avg.pct_of_peak_sustained_active") or num(m, "smsp__issue_active.avg.pct"),
            "achieved_occupancy_pct": num(m, "sm__warps_active.avg.pct_of_peak_sustained_active"),
            "l1_hit_pct": num(m, "l1tex__t_sector_hit_rate.pct"), "l2_hit_pct": num(m, "lts__t_sector_hit_rate.pct"),
            "l1tex_global_ld_sectors": num(m, "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum"),
            "l1tex_global_ld_requests": num(m, "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum"),
            "registers_per_thread": num(m, "launch__registers_per_thread"), "grid": m.get("launch__grid_size"),
            "block": m.get("launch__block_size"),
        }
        base = os.path.join(ROOT, "profiles", "%s_%s" % (tag, name))
        det = subprocess.run(["ncu", "-i", rep, "--page", "details"], capture_output=True, text=True).stdout
        open(base + "_details.txt", "w").write(det)
        src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
        tmp = "/tmp/_src_%s.csv" % name
        open(tmp, "w").write(src)
        st = subprocess.run([sys.executable, os.path.join(ROOT, "profiles", "ncu_top.py"), tmp, "32"], capture_output=True, text=True).stdout
        open(base + "_stalls.txt", "w").write(st)
        # stall mix as numbers (first two lines of ncu_top's output)
        res[name]["stall_mix"] = st.splitlines()[1] if len(st.splitlines()) > 1 else ""
    json.dump(res, open(os.path.join(ROOT, "profiles", "%s_traffic.json" % tag), "w"), indent=1)
    print(json.dumps(res, indent=1))
