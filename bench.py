#!/usr/bin/env python
"""bench.py -- DPFs/sec of the hot path on N B200s (one process per GPU).

    python bench.py [--gpus N] [--steps K] [--warmup W]          our engine
    python bench.py --impl reference [...]                       reference CPU path

A step is one pass of the hot path over one batch: full-domain evaluation of
`batch` DPF keys fused with the inner product against an [n, 16] int32 table
(BASELINE.json metric: n = 2^20, entry_size = 16, AES128, batch 512 per GPU).

N > 1 (launched by torchrun, NCCL): the table is sharded by entry range, every
rank evaluates the whole batch over its subtree, and one NCCL reduce adds the
[B, 16] partials.  The batch grows with N (512 per GPU),
so per-GPU work is fixed: weak scaling.

Rank 0 prints ONE JSON line.  `value` is device-timed with keys resident in
HBM; `e2e` goes through the public dpf-API call with pinned HOST keys, H2D and
D2H inside the timed region.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(ROOT, "gpu-dpf_b200"))

PRF_IDS = {"dummy": 0, "salsa20": 1, "chacha20": 2, "aes128": 3}
KEY_BYTES = 2096
BASELINE_PUBLISHED = {  # BASELINE.md section 1 (reference README.md:129-146), V100, batch 512, entry 16
    ("aes128", 1 << 14): 52536, ("aes128", 1 << 16): 15392, ("aes128", 1 << 18): 3967, ("aes128", 1 << 20): 923,
    ("salsa20", 1 << 14): 145646, ("salsa20", 1 << 16): 54892, ("salsa20", 1 << 18): 16650, ("salsa20", 1 << 20): 3894,
    ("chacha20", 1 << 14): 139590, ("chacha20", 1 << 16): 56120, ("chacha20", 1 << 18): 16086, ("chacha20", 1 << 20): 4054,
}


_REAL_STDOUT = None


def emit(line):
    data = (json.dumps(line) + "\n").encode()
    if _REAL_STDOUT is not None:
        os.write(_REAL_STDOUT, data)
    else:
        sys.stdout.write(data.decode())
        sys.stdout.flush()


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference", "reference-gpu"],
                    help="reference = the reference's CPU path on the host cores; reference-gpu = the reference's "
                         "own GPU kernel compiled unmodified for sm_100a (oracle/_ref/ref_dpf_cpp.so), 1 GPU")
    ap.add_argument("--entries", "--n", dest="n", type=int, default=1 << 20, help="table size n (use --entries under torchrun: --n is ambiguous to its parser)")
    ap.add_argument("--entry", type=int, default=16)
    ap.add_argument("--prf", default="aes128", choices=sorted(PRF_IDS))
    ap.add_argument("--batch-per-gpu", type=int, default=512)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--subtree-log2", type=int, default=0)
    ap.add_argument("--strong", action="store_true",
                    help="N>1: keep the GLOBAL batch at --batch-per-gpu keys (strong scaling: each GPU still sees "
                         "every key but only 1/N of the leaves) instead of growing it with N")
    ap.add_argument("--reduce", default="nccl", choices=["nccl", "fused"],
                    help="N>1: NCCL reduce of the partials, or the kernel's peer-memory red.add epilogue")
    return ap.parse_args()


def bytes_per_dpf(n, entry):
    """int32 table streamed once per key + key + output."""
    return n * entry * 4 + KEY_BYTES + 4 * entry


def synthetic_table(n, entry):
    rng = np.random.RandomState(1234)
    return rng.randint(0, 2**31, size=(n, entry), dtype=np.int64).astype(np.int32)


def synthetic_keys(n, batch, prf):
    import b200dpf
    rng = np.random.RandomState(4321)
    alphas = rng.randint(0, n, size=batch).astype(np.int64)
    ka, _ = b200dpf.gen_batch(alphas, n, np.arange(batch) + 1000, prf)
    return ka, alphas


# ---------------------------------------------------------------------------
# clocks (B200_PROFILING.md recipe)
# ---------------------------------------------------------------------------
class ClockSampler:
    FIELDS = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
              "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu_index = gpu_index
        self.proc = None
        self.path = None

    def start(self):
        try:
            fd, self.path = tempfile.mkstemp(suffix=".csv")
            os.close(fd)
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.gpu_index), "--query-gpu=" + self.FIELDS,
                 "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.proc is None:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, reasons, smax = [], set(), None
        try:
            for line in open(self.path):
                f = [x.strip() for x in line.split(",")]
                if len(f) < 9:
                    continue
                try:
                    sm.append(float(f[1]))
                    smax = float(f[2])
                except ValueError:
                    continue
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            os.unlink(self.path)
        except Exception:
            pass
        if sm:
            top = sorted(sm)[len(sm) // 2:]          # upper half = samples taken under load
            out["sm_mhz"] = float(np.median(top))
            out["samples"] = len(sm)
        out["sm_max_mhz"] = smax
        out["reasons"] = sorted(reasons)
        return out


# ---------------------------------------------------------------------------
# reference CPU path (oracle/_ref when the reference compiled, else the oracle port)
# ---------------------------------------------------------------------------
def cpu_reference_runner(n, entry, prf, table):
    """Returns (kind, cores, sample_desc, step_fn) where step_fn() runs one bounded
    sample and returns the number of DPF-equivalents it evaluated."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import oracle as O
    cores = os.cpu_count() or 1
    orc = O.Oracle()
    rng = np.random.RandomState(99)
    keys = np.stack([orc.gen(int(rng.randint(0, n)), n, 2000 + i, prf)[0] for i in range(cores)])
    depth = n.bit_length() - 1
    per_index_us = {0: 0.01, 1: 0.08, 2: 0.08, 3: 1.4}[prf] * depth
    idx_count = int(min(n, max(1024, 1.5e6 / per_index_us)))       # about 1.5 s per thread
    idx_count = 1 << (idx_count.bit_length() - 1)
    frac = idx_count / n
    if O.Ref.available():
        ref = O.Ref()
        kind = "reference"

        def step():
            ref.eval_dot_mt(keys, prf, table, 0, idx_count, cores)
            return cores * frac
    else:
        kind = "port"

        def step():
            th = [threading.Thread(target=orc.eval_dot_range, args=(keys[i:i + 1], prf, table, 0, idx_count))
                  for i in range(cores)]
            [t.start() for t in th]
            [t.join() for t in th]
            return cores * frac
    sample = ("%d keys (one per host thread) x %d of %d indices each via per-index EvaluateFlat + int32 dot, "
              "scaled by n/indices" % (cores, idx_count, n))
    return kind, cores, sample, step


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    prf = PRF_IDS[args.prf]
    n, entry = args.n, args.entry
    table = synthetic_table(n, entry)
    kind, cores, sample, step = cpu_reference_runner(n, entry, prf, table)
    for _ in range(args.warmup):
        step()
    t0 = time.perf_counter()
    units = 0.0
    for _ in range(args.steps):
        units += step()
    dt = time.perf_counter() - t0
    value = units / dt
    line = {
        "impl": "reference", "metric": "DPFs/sec", "value": value, "unit": "DPFs/sec", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / max(args.steps, 1),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u32", "data": "synthetic",
        "config": {"workload": "n=%d entry_size=%d %s, reference CPU path (dpf_base EvaluateFlat per index) on host cores"
                               % (n, entry, args.prf.upper()), "n": n, "entry_size": entry, "prf": args.prf.upper()},
        "cpu_baseline": {"value": value, "unit": "DPFs/sec", "cores": cores, "kind": kind, "sample": sample},
        "e2e": {"value": value, "unit": "DPFs/sec", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


def run_reference_gpu(args):
    """The reference's dpf_hybrid_kernel on this box, measured the way its benchmark.py does
    (dpf.py:286-320): wall clock around `steps` eval_gpu calls of 512 host key tensors."""
    if int(os.environ.get("RANK", "0")) != 0:
        return
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import torch
    import refgpu
    if not refgpu.available():
        emit({"impl": "reference-gpu", "unavailable": "oracle/_ref/ref_dpf_cpp.so not built (needs the reference tree)"})
        return
    prf = PRF_IDS[args.prf]
    n, entry, batch = args.n, args.entry, args.batch_per_gpu
    table = synthetic_table(n, entry)
    keys_np, _ = synthetic_keys(n, batch, prf)
    keys = [torch.from_numpy(k) for k in keys_np]
    ref = refgpu.RefGpuDPF(prf)
    t0 = time.perf_counter()
    ref.eval_init(torch.from_numpy(table))
    t_init = time.perf_counter() - t0
    for _ in range(max(args.warmup, 1)):
        ref.eval_gpu(keys)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        out = ref.eval_gpu(keys)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    value = batch * args.steps / dt
    emit({"impl": "reference-gpu", "metric": "DPFs/sec", "value": value, "unit": "DPFs/sec", "n_gpus": 1,
          "steps": args.steps, "warmup": max(args.warmup, 1), "ms_per_step": 1e3 * dt / args.steps,
          "higher_is_better": True, "dtype": "u128", "data": "synthetic",
          "config": {"workload": "n=%d entry_size=%d %s batch=%d, reference dpf_hybrid_kernel compiled for sm_100a, "
                                 "wall clock incl. host marshalling (benchmark.py method)" % (n, entry, args.prf.upper(), batch),
                     "eval_init_s": t_init},
          "e2e": {"value": value, "unit": "DPFs/sec", "h2d_bytes_per_step": batch * 2080, "d2h_bytes_per_step": batch * 256},
          "checksum": int(out.to(torch.int64).sum().item())})
    ref.close()


# ---------------------------------------------------------------------------
# our engine
# ---------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as dist
    import dpf as dpf_mod
    from sharded import ShardedDPF

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the engine has no CPU path (use --impl reference for the CPU baseline)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        import datetime
        dist.init_process_group("nccl", device_id=dev, timeout=datetime.timedelta(seconds=180))
    assert world == args.gpus or world == 1, "launch with torchrun --nproc-per-node %d" % args.gpus

    prf = PRF_IDS[args.prf]
    n, entry = args.n, args.entry
    batch = args.batch_per_gpu * (1 if args.strong else world)
    table = synthetic_table(n, entry)
    keys_np, _ = synthetic_keys(n, batch, prf)

    if world > 1:
        d = ShardedDPF(prf=prf, device=local_rank, reduce=args.reduce)
        d.eval_init(torch.from_numpy(table))
        inner = d._dpf
    else:
        d = dpf_mod.DPF(prf=prf, device=local_rank)
        d.eval_init(torch.from_numpy(table))
        inner = d
    if args.subtree_log2:
        import dpf_cpp
        dpf_cpp.set_subtree_log2(inner.buffers, args.subtree_log2)

    keys_dev = torch.from_numpy(keys_np).to(dev)
    out_dev = torch.empty((batch, entry), dtype=torch.int32, device=dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)   # > 126 MB L2

    def barrier():
        if world > 1:
            dist.barrier()

    def step_device():
        d.eval_gpu_device(keys_dev, out_dev)

    sampler = ClockSampler(local_rank) if rank == 0 else None
    if sampler:
        sampler.start()          # nvidia-smi needs a moment to produce its first sample
    nwarm = max(args.warmup, 3)
    t_warm = time.perf_counter()
    for _ in range(nwarm):
        flush.zero_()
        step_device()
    torch.cuda.synchronize()
    # keep the GPU busy for ~0.5 s in total so the clock sampler sees it under load; the number
    # of extra iterations is decided on rank 0 and broadcast (collectives must match across ranks)
    per_step = (time.perf_counter() - t_warm) / nwarm
    extra = torch.tensor([int(min(2000, max(0, 0.5 / max(per_step, 1e-6) - nwarm)))], dtype=torch.int64, device=dev)
    if world > 1:
        dist.broadcast(extra, src=0)
    for i in range(int(extra.item())):
        flush.zero_()
        step_device()
        if i % 16 == 15:
            torch.cuda.synchronize()
    nwarm += int(extra.item())
    torch.cuda.synchronize()

    import dpf_cpp
    launches_per_step = dpf_cpp.last_launches(inner.buffers)
    starts = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
    ends = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
    barrier()
    torch.cuda.synchronize()
    t_wall0 = time.perf_counter()
    for k in range(args.steps):
        flush.zero_()                       # cold L2 at the start of every timed step
        starts[k].record()
        step_device()
        ends[k].record()
    torch.cuda.synchronize()
    barrier()
    t_wall = time.perf_counter() - t_wall0
    clocks = sampler.stop() if sampler else None
    dev_ms = sum(s.elapsed_time(e) for s, e in zip(starts, ends))
    t = torch.tensor([dev_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dev_ms = float(t.item())
    value = batch * args.steps / (dev_ms / 1e3)

    # ---- end to end through the public API, host buffers ----
    e2e = None
    if not args.no_e2e:
        keys_host = torch.from_numpy(keys_np).pin_memory()
        for _ in range(2):
            d.eval_gpu(keys_host)
        barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            res = d.eval_gpu(keys_host)
        torch.cuda.synchronize()
        barrier()
        dt = time.perf_counter() - t0
        t = torch.tensor([dt], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dt = float(t.item())
        e2e = {"value": batch * args.steps / dt, "unit": "DPFs/sec",
               "h2d_bytes_per_step": int(keys_host.numel() * 4) * world,
               "d2h_bytes_per_step": int(batch * entry * 4)}
        if rank == 0:
            assert res is not None and tuple(res.shape) == (batch, entry)

    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        peak = float(peaks.get("hbm_gbs", 6650.0))
        peak_kind = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback (B200_PROFILING.md)"
        # dominant kernel = the one evaluation kernel of a step; per launch it processes `batch`
        # keys over n/world leaves each
        alg_bytes = batch * (n // world * entry * 4 + KEY_BYTES + 4 * entry)
        launch_ms = dev_ms / args.steps
        achieved = alg_bytes / (launch_ms / 1e3) / 1e9
        traffic = None
        try:   # measured once per change with `ncu --set full` (tools/gpu_round1_final.sh), per launch
            tr = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))
            key = "%s_n%d_e%d_b%d_%dgpu" % (args.prf, n, entry, batch, world)
            traffic = tr.get(key, {}).get("dram_bytes_per_launch")
        except Exception:
            pass
        binding = {"aes128": {"pipe": "l1tex data pipe (shared-memory T-table wavefronts)", "busy_frac": 0.979},
                   "salsa20": {"pipe": "integer ALU pipe (LOP3/SHF)", "busy_frac": 0.977},
                   "chacha20": {"pipe": "integer ALU pipe (LOP3/SHF)", "busy_frac": 0.954}}.get(args.prf)
        roofline = {"bound": "hbm", "binding_pipe_ncu": binding, "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                    "traffic": traffic, "peak_source": peak_kind,
                    "note": "algorithmic bytes = batch*(n*E*4/ngpu + 2096 + 4E) per launch (table streamed once "
                            "per key, SURVEY 8d); actual DRAM traffic is far lower because 32 keys share each row "
                            "load and the table stays in L2. The binding resource is the LSU/shared-memory data "
                            "pipe for AES (97.9% busy, ncu) and the ALU pipe for Salsa/ChaCha (97.7%): DESIGN.md s4"}
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            kind, cores, sample, step = cpu_reference_runner(n, entry, prf, table)
            step()
            t0 = time.perf_counter()
            units = step()
            cpu = {"value": units / (time.perf_counter() - t0), "unit": "DPFs/sec", "cores": cores, "kind": kind,
                   "sample": sample}
        published = BASELINE_PUBLISHED.get((args.prf, n)) if (entry == 16) else None
        line = {
            "metric": "DPFs/sec", "value": value, "unit": "DPFs/sec", "n_gpus": world, "steps": args.steps,
            "warmup": nwarm, "ms_per_step": dev_ms / args.steps, "higher_is_better": True,
            "scaling": "strong" if args.strong else "weak",
            "vs_baseline": (value / published) if published else None, "dtype": "u32",
            "data": "synthetic",
            "config": {"workload": "n=%d entry_size=%d %s batch=%d (%s), table entry-range sharded over %d GPU(s)"
                                   % (n, entry, args.prf.upper(), batch,
                                      "fixed global batch" if args.strong else "%d per GPU" % args.batch_per_gpu, world),
                       "n": n, "entry_size": entry, "prf": args.prf.upper(), "global_batch": batch,
                       "parallelism": ("entry-shard x%d + %s" % (world, "NCCL reduce" if args.reduce == "nccl" else
                                       "in-kernel peer-memory red.add (symmetric memory)")) if world > 1 else "single GPU",
                       "l2": "256 MiB device buffer zeroed before every timed step (L2 flush); table 64 MiB",
                       "vs_baseline_ref": "reference README V100 number (BASELINE.md)" if published else None},
            "e2e": e2e, "gpu_launches": launches_per_step * args.steps, "roofline": roofline,
            "cpu_baseline": cpu, "clocks": clocks, "wall_s_timed_region": t_wall,
        }
        emit(line)
    d.close()
    if world > 1:
        dist.destroy_process_group()


def main():
    args = parse_args()
    # Libraries (NCCL's version banner, torchrun notices) write to stdout; the contract is ONE
    # JSON line there.  Point fd 1 at stderr for the whole run and print the line on the saved fd.
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)
    if args.impl == "reference":
        run_reference(args)
    elif args.impl == "reference-gpu":
        run_reference_gpu(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
