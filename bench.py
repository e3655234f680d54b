#!/usr/bin/env python3
"""bench.py -- fold-stage throughput on B200 (BASELINE.json metric: folded nt/sec, RNALfold -L 300).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A "step" is one pass of the hot path (encode -> c/fML band fill -> f3 -> plan/traceback/emission) over
one batch of synthetic loci.  Workload = BASELINE.json configs[1] ("parity-10k": 10 000 loci, 300-600 nt,
GC 0.40 with embedded hairpins, L=300), one such batch PER GPU (weak scaling, loci are
independent: no data-path collective).  `value` = nt/s with the raw sequences already resident in HBM
and results left in HBM (CUDA events on the launching stream, max over ranks); `e2e` = the same metric
through the public host API (host buffers in, hit records out, copies inside the timed region).
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
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
from corpus import synth_loci  # noqa: E402  (pure generator, no oracle code)

METRIC = "folded nt/sec (RNALfold -L 300, bit-exact MFE)"
UNIT = "nt/s"
SPAN = 300


WORKLOADS = {   # name -> (length law, base seed, BASELINE configs index)
    "parity": ("parity", 1001, 1), "arabidopsis": ("arabidopsis", 1002, 2), "long": ("long", 1003, 3), "sweep": ("sweep", 1004, 4)}
WORKLOAD = "parity"


def workload(rank, nloci):
    law, seed, _ = WORKLOADS[WORKLOAD]
    return synth_loci(seed + rank, nloci, law)


# ---------------------------------------------------------------------------------- work model
def g_of_d():
    G = np.zeros(4200, np.int64)
    for d in range(6, len(G)):
        K = min(30, d - 6)
        G[d] = sum(min(30 - u, d - u - 6) + 1 for u in range(K + 1))
    return G


def typed_fraction(seqs, span, sample=48):
    """rho = fraction of band cells (i,j) whose bases can pair, on a seeded sample."""
    pair = np.zeros((8, 8), bool)
    for a, b in ((2, 3), (3, 2), (3, 4), (4, 3), (1, 4), (4, 1)):
        pair[a, b] = True
    code = np.zeros(256, np.int64)
    for ch, v in zip("ACGUT", (1, 2, 3, 4, 4)):
        code[ord(ch)] = v
        code[ord(ch.lower())] = v
    typed = cells = 0
    for s in seqs[:sample]:
        c = code[np.frombuffer(s.encode(), np.uint8)]
        n = len(c)
        for d in range(4, min(span, n - 1) + 1):
            cells += n - d
            if d < min(span, n):
                typed += int(pair[c[:n - d], c[d:]].sum())
    return typed / max(cells, 1)


def algorithmic_terms(lens, span, rho):
    """T_alg of SURVEY.md 8(d): sum over cells of max(0,d-8) + rho^2*G(d) + 16 min-plus terms."""
    G = g_of_d()
    tot = 0.0
    cells = 0
    for n in lens:
        dmax = min(span, n - 1)
        if dmax < 4:
            continue
        d = np.arange(4, dmax + 1)
        cnt = n - d
        tot += float((cnt * (np.maximum(0, d - 8) + rho * rho * G[d] + 16)).sum())
        cells += int(cnt.sum())
    return tot, cells


# ---------------------------------------------------------------------------------- clocks
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "50"], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")] + [time.time()])

    def window(self, t0, t1):
        """keep the samples taken inside [t0, t1] (the sampler is started early: nvidia-smi takes ~0.1 s to come up)"""
        inside = [r for r in self.rows if t0 <= r[-1] <= t1 + 0.05]
        if len(inside) >= 1:
            self.rows = inside

    def stop(self):
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=5)
            except Exception:
                pass
        self.rows = [r[:-1] for r in self.rows]
        sm = [float(r[1]) for r in self.rows if len(r) >= 9 and r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in self.rows if len(r) >= 9 and r[2].replace(".", "").isdigit()]
        reasons = set()
        for r in self.rows:
            if len(r) >= 9:
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ---------------------------------------------------------------------------------- CPU reference arm
def _rlf_binary():
    rlf = os.path.join(ROOT, "oracle", "_ref", "RNALfold")
    if os.path.exists(rlf) and os.access(rlf, os.X_OK):
        return rlf, "reference"
    cli = os.path.join(ROOT, "oracle", "_build", "lfold_oracle")
    if not os.path.exists(cli):
        subprocess.run(["make", "-C", os.path.join(ROOT, "oracle"), "all"], check=True, stdout=subprocess.DEVNULL)
    return cli, "port"


def cpu_fold_stage(seqs, span, cores):
    """The reference's fold stage on host cores: P contiguous shards by locus count
    (num_each_piece = int(total/P)+1, miR_PREFeR.py:1329-1354), one `RNALfold -L span` process per
    shard, all concurrent (MP:3113-3118), output to tmpfs.  Returns (seconds, kind)."""
    binary, kind = _rlf_binary()
    P = max(1, min(cores, len(seqs)))
    per = len(seqs) // P + 1
    tmp = tempfile.mkdtemp(dir="/dev/shm" if os.path.isdir("/dev/shm") else None)
    files = []
    for k in range(P):
        part = seqs[k * per:(k + 1) * per]
        if not part:
            continue
        fn = os.path.join(tmp, "in_%d.fa" % k)
        with open(fn, "w") as f:
            for r, s in enumerate(part):
                f.write(">r%d_%d\n%s\n" % (k, r, s))
        files.append(fn)
    t0 = time.perf_counter()
    procs = []
    for fn in files:
        procs.append(subprocess.Popen([binary, "-L", str(span)], stdin=open(fn), stdout=open(fn + ".out", "w")))
    for p in procs:
        if p.wait() != 0:
            raise RuntimeError("reference fold process failed")
    dt = time.perf_counter() - t0
    for fn in files:
        os.remove(fn)
        os.remove(fn + ".out")
    os.rmdir(tmp)
    return dt, kind, len(files)


def run_reference(args, rank, world):
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    # bounded sample: about 0.12 s of RNALfold per locus, one shard per core; keep the whole run near 100 s
    per_core = int(max(8, min(args.ref_loci_per_core, 100.0 / (args.steps + 1) / 0.12)))
    per_step = max(cores, min(args.loci, cores * per_core))
    seqs = workload(0, per_step)
    nt = sum(len(s) for s in seqs)
    for _ in range(min(args.warmup, 1)):
        cpu_fold_stage(seqs[:cores], SPAN, cores)
    t = 0.0
    kind, used = "reference", cores
    for _ in range(args.steps):
        dt, kind, used = cpu_fold_stage(seqs, SPAN, cores)
        t += dt
    val = nt * args.steps / t
    sample = "%d loci (%d nt) of parity-10k seed 1001 per step, %d concurrent RNALfold processes" % (per_step, nt, used)
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * t / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "int32", "data": "synthetic",
        "config": {"workload": "parity-10k (BASELINE configs[1]): loci 300-600 nt, L=300; bounded sample per step",
                   "span_L": SPAN, "loci_per_step": per_step},
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": used, "kind": kind, "sample": sample},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0}))


# ---------------------------------------------------------------------------------- ours
def run_ours(args, rank, world, local_rank):
    import torch
    import torch.distributed as dist
    import mir_prefer_b200 as mp
    from mir_prefer_b200 import _lib
    if not os.path.exists(_lib.LIB_PATH):
        raise RuntimeError("libmirfold.so missing (run python __graft_entry__.py); there is no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    seqs = workload(rank, args.loci)
    lens = [len(s) for s in seqs]
    nt = sum(lens)
    mf = mp.MirFold(devices=[local_rank])
    buf, off = mf.pack(seqs)
    pinned = torch.from_numpy(buf.copy()).pin_memory()
    d_buf = pinned.cuda(non_blocking=False)
    stream = torch.cuda.current_stream().cuda_stream

    # ---- kernel-resident timing (value): inputs in HBM, results stay in HBM
    fill_ms, dev_ms, launches, tracebacks, cells = [], [], 0, 0, 0
    stage = {"ms_fill": 0.0, "ms_f3": 0.0, "ms_trace": 0.0}
    sampler = ClockSampler(local_rank)
    sampler.start()
    for _ in range(args.warmup):
        mf.fold_device(d_buf.data_ptr(), off, SPAN, stream=stream).close()
    barrier()
    t_w0 = time.time()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        r = mf.fold_device(d_buf.data_ptr(), off, SPAN, stream=stream)
        fill_ms.append(r.stats["ms_fill"]); dev_ms.append(r.stats["ms_device"])
        for k in stage:
            stage[k] += r.stats[k] / args.steps
        n_chunks = r.stats["n_chunks"]
        launches += r.stats["kernel_launches"]; tracebacks = r.stats["tracebacks"]; cells = r.stats["cells"]
        r.close()
    e1.record()
    barrier()
    sampler.window(t_w0, time.time())
    clocks = sampler.stop()
    t_dev = e0.elapsed_time(e1) * 1e-3
    # ---- end to end through the public API: host buffers in, hit records out
    host_buf = pinned.numpy()
    for _ in range(min(args.warmup, 2)):
        mf.fold_packed(host_buf, off, SPAN).close()
    barrier()
    t0 = time.perf_counter()
    h2d = d2h = 0
    for _ in range(args.steps):
        r = mf.fold_packed(host_buf, off, SPAN)
        h2d, d2h = r.stats["h2d_bytes"], r.stats["d2h_bytes"]
        nhits = r.nhits
        e2e_split = {k: r.stats[k] for k in ("ms_total", "ms_h2d", "ms_device", "ms_d2h")}
        r.close()
    barrier()
    t_e2e = time.perf_counter() - t0

    tt = torch.tensor([t_dev, t_e2e], dtype=torch.float64, device="cuda")
    tot = torch.tensor([float(nt), float(cells)], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        dist.all_reduce(tot, op=dist.ReduceOp.SUM)
    t_dev, t_e2e = tt.tolist()
    nt_all, cells_all = tot.tolist()

    if rank == 0:
        rho = typed_fraction(seqs, SPAN)
        t_alg, _ = algorithmic_terms(lens, SPAN, rho)
        peak_a, peak_dpx = mf.int_peak()
        peak = max(peak_a, peak_dpx)
        fill_s = float(np.mean(fill_ms)) * 1e-3
        achieved = t_alg / fill_s
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
            hbm_peak, hbm_src = float(peaks["hbm_gbs"]), "measured"
        except Exception:
            hbm_peak, hbm_src = 6650.0, "fallback"
        band_bytes = cells * 8.0  # c + fML written once, int32
        traffic_note = None
        traffic = None            # dram read+write of the dominant fill launch, from the committed ncu --set full capture
        try:
            tr = json.load(open(os.path.join(ROOT, "profiles", "fill_traffic.json")))
            if WORKLOAD == "parity" and args.loci == 10000 and SPAN == 300:
                traffic = tr["dram_bytes_read"] + tr["dram_bytes_write"]      # bytes per launch
                traffic_note = "%s, %s" % (tr["kernel"], tr["source"])
        except Exception:
            pass
        out = {
            "metric": METRIC, "value": nt_all * args.steps / t_dev, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * t_dev / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "int32", "data": "synthetic",
            "config": {"workload": ("parity-10k (BASELINE configs[1]): %d loci/GPU, 300-600 nt, GC 0.40, embedded hairpins, L=300"
                                    % args.loci) if WORKLOAD == "parity" and SPAN == 300 else
                                   "%s law (BASELINE configs[%d]), %d loci/GPU, L=%d" % (WORKLOAD, WORKLOADS[WORKLOAD][2], args.loci, SPAN),
                       "span_L": SPAN, "loci_per_gpu": args.loci, "nt_per_gpu": nt,
                       "dp_cells_per_gpu": int(cells), "cache": "inputs+band workspace (%.1f GB) far larger than L2; no flush needed"
                                   % (band_bytes / 1e9), "tracebacks_per_step": int(tracebacks)},
            "dp_cells_per_s": cells_all * args.steps / t_dev,
            "e2e": {"value": nt_all * args.steps / t_e2e, "unit": UNIT, "h2d_bytes_per_step": int(h2d),
                    "d2h_bytes_per_step": int(d2h), "ms_per_step": 1e3 * t_e2e / args.steps, "hits_per_step": int(nhits), "last_step_split": e2e_split},
            "gpu_launches": int(launches),
            "stage_ms": dict(stage, chunks=int(n_chunks)),
            "clocks": clocks,
            "roofline": {"bound": "int32-issue", "kernel": "k_fill", "achieved": achieved / 1e12, "peak": peak / 1e12,
                         "unit": "Tterm/s (1 min-plus term = 1 add + 1 min)", "frac": achieved / peak,
                         "peak_source": "measured live: VIADDMNMX stream (mirfold_int_peak), add+min %.2f / DPX %.2f Tterm/s"
                                        % (peak_a / 1e12, peak_dpx / 1e12),
                         "algorithmic_terms_per_launch": t_alg, "terms_per_cell": t_alg / max(cells, 1), "rho": rho,
                         "kernel_ms": fill_s * 1e3, "kernel_share_of_step": fill_s / (t_dev / args.steps),
                         "traffic": traffic, "traffic_note": traffic_note,
                         "hbm": {"bound": "hbm", "achieved": band_bytes / fill_s / 1e9, "peak": hbm_peak, "unit": "GB/s",
                                 "frac": band_bytes / fill_s / 1e9 / hbm_peak, "peak_source": hbm_src,
                                 "note": "band store of c+fML (8 B/cell) only; the kernel is integer-issue bound"}},
        }
        if world == 1:
            # the reference-compatible boundaries on the same batch (informational, outside the timed regions):
            # RNALfold text in -> byte-identical text out, and the stage-1 candidate-structure tuples
            text = "".join(">locus%d:%d-%d + 1-22 0 1,22,+\n%s\n" % (k, 1, len(x) + 1, x) for k, x in enumerate(seqs))
            mf.fold_text_bytes(text, SPAN)
            t0 = time.perf_counter()
            nbytes = len(mf.fold_text_bytes(text, SPAN))
            t_text = time.perf_counter() - t0
            with mf.fold_packed(host_buf, off, SPAN) as r:
                t0 = time.perf_counter()
                per_rec = r.classify(55)
                nstruct = sum(len(x) for x in per_rec)
                t_cls = time.perf_counter() - t0
            # stage 3 on the same structures: one synthetic 21-nt mature on the 5' arm of every structure
            queries = [(ss, (fs + 8, fs + 29), fs, 1, lens[k] + 1, "+") for k, recs_ in enumerate(per_rec) for (_, fs, ss, _) in recs_]
            t0 = time.perf_counter()
            verdicts = mf.duplex(queries)
            t_dup = time.perf_counter() - t0
            out["drop_in"] = {"rnalfold_text_in_out_ms": 1e3 * t_text, "text_bytes": nbytes,
                              "classify_structures_ms": 1e3 * t_cls, "structures": nstruct,
                              "duplex_queries_ms": 1e3 * t_dup, "duplex_queries": len(queries),
                              "duplex_pass": sum(1 for v in verdicts if not isinstance(v, str))}
        if world == 1 and not args.no_cpu:
            cores = os.cpu_count() or 1
            sample_n = max(cores, min(args.loci, cores * args.ref_loci_per_core))
            sample = seqs[:sample_n]
            dt, kind, used = cpu_fold_stage(sample, SPAN, cores)
            snt = sum(len(s) for s in sample)
            one = sample[:12]                                   # also the P=1 figure
            dt1, _, _ = cpu_fold_stage(one, SPAN, 1)
            out["cpu_baseline"] = {"value": snt / dt, "unit": UNIT, "cores": used, "kind": kind,
                                   "sample": "first %d loci (%d nt) of the same workload, %d concurrent RNALfold -L 300 "
                                             "processes, %.1f s" % (sample_n, snt, used, dt),
                                   "single_core_value": sum(len(x) for x in one) / dt1,
                                   "dp_cells_per_s": snt / dt * (cells / max(nt, 1))}
        print(json.dumps(out))
    mf.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--loci", type=int, default=10000, help="loci per GPU (BASELINE configs[1]: 10000)")
    ap.add_argument("--ref-loci-per-core", type=int, default=96, help="CPU legs: loci per host core and step (about 10 s of RNALfold)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--workload", default="parity", choices=sorted(WORKLOADS),
                    help="length law of SURVEY 8(d); the bench line of record is the default (parity-10k)")
    ap.add_argument("--span", type=int, default=300, help="RNALfold -L (default 300)")
    args = ap.parse_args()
    global WORKLOAD, SPAN
    WORKLOAD, SPAN = args.workload, args.span
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    if world == 1 and args.gpus > 1:
        # convenience: re-launch under torchrun
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(args.gpus),
               "--master-addr", "127.0.0.1", "--master-port", "29513", os.path.abspath(__file__)] + sys.argv[1:]
        sys.exit(subprocess.call(cmd))
    run_ours(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
