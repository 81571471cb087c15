#!/usr/bin/env python
"""Turn the raw measurements of one GPU call into the two small files bench.py's pipe roofline reads:

  profiles/int_peaks.json   measured issue rates per SM (tools/microbench/int_pipes.cu output)
  profiles/pipe_ops.json    measured warp-instructions per node pair of the evaluation kernel, per PRF
                            (ncu sm__inst_executed_pipe_* / l1tex__data_pipe_lsu_wavefronts of one launch)

usage: python tools/make_pipe_profiles.py gpurun_out/r2_int_pipes.jsonl gpurun_out/r2b_pipes_{prf}_n1048576.csv
"""
import csv
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def read_ncu_metrics(path):
    rows = [r for r in csv.reader(open(path)) if len(r) > 14 and r[0] != "ID"]
    return {r[12]: float(r[14].replace(",", "")) for r in rows}, (rows[0][4] if rows else "")


def main():
    pipes_jsonl = sys.argv[1]
    pattern = sys.argv[2]
    dev, pipe, qr = {}, {}, []
    for line in open(pipes_jsonl):
        d = json.loads(line)
        if d["kind"] == "device":
            dev = d
        elif d["kind"] == "pipe":
            pipe[d["name"]] = d
        elif d["kind"] == "quarter_round":
            qr.append({k: d[k] for k in ("cipher", "fma_rot_mask", "clk_per_qr_sm")})
    peaks = {
        "source": "tools/microbench/int_pipes.cu on %s (%d SMs), warp-instructions per clock per SM from the SM cycle "
                  "counter; 16 warps/SM, 8 independent chains per thread" % (dev.get("name"), dev.get("sms", 148)),
        "sms": dev.get("sms", 148), "sm_max_mhz": dev.get("clock_mhz_nominal", 1965.0),
        "peak_warp_inst_per_clk_sm": {
            "alu": max(pipe[n]["alu_inst_per_clk_sm"] for n in ("LOP3", "SHF", "PRMT", "IADD3")),
            "fma": pipe["IMAD"]["fma_inst_per_clk_sm"],
            "lsu": pipe["PRMT+LDS.32 conflict-free"]["lsu_inst_per_clk_sm"],
        },
        "per_instruction": {n: {"alu": p["alu_inst_per_clk_sm"], "fma": p["fma_inst_per_clk_sm"], "lsu": p["lsu_inst_per_clk_sm"],
                                "steps_per_clk_sm": p["steps_per_clk_sm"]} for n, p in pipe.items()},
        "notes": {
            "IMAD.WIDE": "half rate: %.2f/clk/SM (IMAD %.2f)" % (pipe["IMAD.WIDE+LOP3"]["fma_inst_per_clk_sm"], pipe["IMAD"]["fma_inst_per_clk_sm"]),
            "dual_issue": "LOP3+IMAD interleaved issues %.2f + %.2f = %.2f warp-inst/clk/SM: the alu and fma pipes overlap, "
                          "but total issue saturates near 3.2/clk/SM, below 2 + 2" % (
                              pipe["LOP3+IMAD"]["alu_inst_per_clk_sm"], pipe["LOP3+IMAD"]["fma_inst_per_clk_sm"],
                              pipe["LOP3+IMAD"]["alu_inst_per_clk_sm"] + pipe["LOP3+IMAD"]["fma_inst_per_clk_sm"]),
        },
        "quarter_rounds_clk_per_qr_sm": qr,
    }
    json.dump(peaks, open(os.path.join(ROOT, "profiles", "int_peaks.json"), "w"), indent=1)
    ops = {}
    for prf, binding, name in (("aes128", "lsu", "LSU / shared-memory data pipe (T-table wavefronts)"),
                               ("salsa20", "alu", "integer ALU pipe (LOP3/SHF)"),
                               ("chacha20", "alu", "integer ALU pipe (LOP3/SHF/PRMT)")):
        path = pattern.replace("{prf}", prf)
        if not os.path.exists(path):
            continue
        m, kname = read_ncu_metrics(path)
        n, batch = 1 << 20, 512
        pairs = (batch // 32) * (n - 1)
        cyc = m["sm__cycles_elapsed.max"]
        sms = peaks["sms"]
        ops[prf] = {
            "binding_pipe": binding, "pipe_name": name, "kernel": kname,
            "source": "ncu --metrics sm__inst_executed_pipe_*.sum,l1tex__data_pipe_lsu_wavefronts.sum, one launch, n=2^20 B=512 "
                      "(%s), divided by (B/32)*(n-1) node pairs" % os.path.basename(path),
            "warp_inst_per_node_pair": {
                "alu": m["sm__inst_executed_pipe_alu.sum"] / pairs,
                "fma": m["sm__inst_executed_pipe_fma.sum"] / pairs,
                "lsu": m["l1tex__data_pipe_lsu_wavefronts.sum"] / pairs,     # wavefronts: what the data pipe spends cycles on
                "lsu_instructions": m["sm__inst_executed_pipe_lsu.sum"] / pairs,
                "all": m["sm__inst_executed.sum"] / pairs,
            },
            "measured_launch": {
                "gpu_time_ms": m["gpu__time_duration.sum"] / 1e6, "sm_cycles_elapsed_max": cyc,
                "alu_inst_per_clk_sm": m["sm__inst_executed_pipe_alu.sum"] / cyc / sms,
                "fma_inst_per_clk_sm": m["sm__inst_executed_pipe_fma.sum"] / cyc / sms,
                "lsu_wavefronts_per_clk_sm": m["l1tex__data_pipe_lsu_wavefronts.sum"] / cyc / sms,
                "all_inst_per_clk_sm": m["sm__inst_executed.sum"] / cyc / sms,
                "dram_bytes": m["dram__bytes_read.sum"] + m["dram__bytes_write.sum"],
            },
        }
        b = ops[prf]["measured_launch"]
        pk = peaks["peak_warp_inst_per_clk_sm"][binding]
        b["binding_pipe_frac_of_measured_peak"] = (b["lsu_wavefronts_per_clk_sm"] if binding == "lsu" else b["alu_inst_per_clk_sm"]) / pk
    json.dump(ops, open(os.path.join(ROOT, "profiles", "pipe_ops.json"), "w"), indent=1)
    for prf, o in ops.items():
        print(prf, {k: round(v, 2) for k, v in o["warp_inst_per_node_pair"].items()}, "binding frac",
              round(o["measured_launch"]["binding_pipe_frac_of_measured_peak"], 4))
    print(json.dumps(peaks["peak_warp_inst_per_clk_sm"]))


if __name__ == "__main__":
    main()
