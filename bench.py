#!/usr/bin/env python
"""bench.py -- aligned pairs/s of the allwave alignment hot path on B200 (BASELINE.json metric).

Workload (config.workload): BASELINE config 2 -- 1,000 x 10 kb synthetic sequences at 5 %
per-haplotype divergence, `-p none` (999,000 directed pairs), scores 0,5,8,2,24,1, mash
orientation.  A "step" is one pass of the whole hot path (orientation -> biWFA -> CIGAR ->
PAF text) over one batch of `--batch` pairs per GPU drawn from that pair list; every rank
aligns its own shard (weak scaling, no collective on the data path).

  value     device-resident: sequences + pair list already in HBM, kernels launched on the
            timed stream, PAF text left in HBM.
  e2e       through the host-facing C-ABI call (aw_load_sequences + aw_align_pairs) with HOST
            buffers: H2D of sequences and pairs and D2H of every PAF line inside the timed region.
  roofline  the alignment kernel alone (CUDA events around it on the launch stream) against the
            measured HBM copy bandwidth, with SURVEY 8(d)'s algorithmic bytes per pair.
  cpu_baseline / --impl reference
            the CPU restatement of the allwave/WFA2 path (oracle/, "port": the Rust reference
            cannot be built in this image), all host threads, on a bounded sample of the same pairs.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

WORKLOAD = "C2: 1000 x 10 kb, 5% divergence per haplotype, -p none (999,000 directed pairs), -s 0,5,8,2,24,1, mash orientation"
SCORES = (0, 5, 8, 2, 24, 1)


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=0, help="pairs per GPU per step (0 = default)")
    ap.add_argument("--nseq", type=int, default=1000)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--e2e-steps", type=int, default=0, help="0 = min(steps, 2)")
    return ap.parse_args()


def c2_sequences(nseq):
    from allwave_b200 import synth

    c, ids, seqs, _ = synth.config("C2", n=nseq)
    return ids, seqs


def job_pairs(nseq, total):
    """deterministic sample of the `-p none` pair list (src/iterator.rs:40-43): every stride-th pair"""
    npairs_all = nseq * (nseq - 1)
    total = min(total, npairs_all)
    stride = max(1, npairs_all // total)
    out = []
    for t in range(total):
        idx = (t * stride) % npairs_all
        i, r = divmod(idx, nseq - 1)
        j = r if r < i else r + 1
        out.append((i, j))
    return out


def algorithmic_bytes(seqs, pairs, paf_bytes):
    # ceil(n/4) + ceil(m/4) + 16 (pair descriptor) + len(PAF line) per pair
    b = paf_bytes
    for q, t in pairs:
        b += (len(seqs[q]) + 3) // 4 + (len(seqs[t]) + 3) // 4 + 16
    return b


class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(gpu_index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200"],
                                      stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.p is None:
            return out
        time.sleep(0.25)
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        self.f.seek(0)
        sm, mx, reasons = [], [], set()
        for line in self.f.read().splitlines():
            parts = [x.strip() for x in line.split(",")]
            if len(parts) < 9:
                continue
            try:
                sm.append(float(parts[1]))
                mx.append(float(parts[2]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), parts[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        try:
            os.unlink(self.f.name)
        except OSError:
            pass
        if sm:
            out.update(sm_mhz=statistics.median(sm), sm_max_mhz=max(mx), reasons=sorted(reasons), samples=len(sm))
        return out


def run_reference(args, rank, world):
    """--impl reference: the CPU restatement (oracle/) on all host threads; rank 0 only."""
    if rank != 0:
        return
    import oracle_lib as O

    cores = os.cpu_count() or 1
    ids, seqs = c2_sequences(args.nseq)
    per_step = max(2, 2 * cores)
    pairs = job_pairs(args.nseq, per_step)
    p = O.params(*SCORES)
    for _ in range(args.warmup):
        O.run_pairs(ids, seqs, pairs[: max(1, cores // 2)], p, use_mash=True, threads=cores)
    t0 = time.perf_counter()
    block = 0
    for _ in range(args.steps):
        r = O.run_pairs(ids, seqs, pairs, p, use_mash=True, threads=cores)
        block += r["sum_block_len"]
    dt = time.perf_counter() - t0
    value = per_step * args.steps / dt
    line = {
        "impl": "reference", "metric": "aligned pairs/s", "value": value, "unit": "pairs/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "int32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "pairs_per_step": per_step, "note": "CPU restatement of the allwave/WFA2 path (oracle/); the Rust reference cannot be built here"},
        "gbp_per_s": block / dt / 1e9,
        "cpu_baseline": {"value": value, "unit": "pairs/s", "cores": cores, "kind": "port", "sample": f"{per_step} pairs of the C2 pair list per step x {args.steps} steps"},
        "e2e": {"value": value, "unit": "pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def main():
    args = parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    import torch.distributed as dist

    import allwave_b200 as aw
    from allwave_b200 import partition

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a B200: the product path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    n_gpus = world
    B = args.batch or 9472  # 16 pairs per resident CTA (148 SMs x 4 CTAs): long enough that the last wave's tail is small
    ids, seqs = c2_sequences(args.nseq)
    # whole job = world*B pairs; host-side greedy (LPT) partition by predicted cost, no collective
    job = job_pairs(args.nseq, B * world)
    shards = partition.partition_pairs(job, [len(s) for s in seqs], world)
    pairs = shards[rank]
    params = aw.make_params(*SCORES)

    ctx = aw.Context(local_rank)
    ctx.load_sequences(ids, seqs)
    batch = aw.Batch(ctx, params, pairs, orientation=aw.AW_ORIENT_MASH, flags=0)
    stream = torch.cuda.current_stream()
    sh = stream.cuda_stream

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- warm-up (also builds the sketches once: they stay resident like the sequences) ----
    for _ in range(args.warmup):
        batch.launch(sh)
    torch.cuda.synchronize()
    batch.fetch(collect=False)
    st = batch.stats()
    if st["failed_pairs"]:
        raise SystemExit(f"{st['failed_pairs']} pairs failed on the GPU path")
    launches_per_step = 2  # orientation kernel + alignment kernel (memsets are not kernels of ours)

    # ---- timed: device-resident ----
    sampler = ClockSampler(local_rank)
    barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    kernel_ms = []
    ev0.record(stream)
    for _ in range(args.steps):
        batch.launch(sh)
        kernel_ms.append(batch.kernel_ms())
    ev1.record(stream)
    barrier()
    ms = ev0.elapsed_time(ev1)
    clocks = sampler.stop()
    batch.fetch(collect=False)
    st = batch.stats()

    # ---- timed: end to end through the host-facing C-ABI call, host buffers ----
    e2e_steps = args.e2e_steps or min(args.steps, 2)
    acc = {"paf_bytes": 0, "n": 0}
    import ctypes as C

    def _cb(rp, _u):
        acc["paf_bytes"] += rp.contents.paf_len
        acc["n"] += 1
        return 0

    cb = aw._cabi.RESULT_CB(_cb)
    arr = aw._cabi.make_pairs(pairs)
    L = aw._cabi.lib()
    ctx2 = aw.Context(local_rank)
    n = len(seqs)
    sa = (C.c_char_p * n)(*seqs)
    la = (C.c_uint64 * n)(*[len(s) for s in seqs])
    ia = (C.c_char_p * n)(*[i.encode() for i in ids])

    def e2e_step():
        t_a = time.perf_counter()
        aw._cabi.check(L.aw_load_sequences(ctx2._h, n, sa, la, ia), "aw_load_sequences")
        t_b = time.perf_counter()
        aw._cabi.check(L.aw_align_pairs(ctx2._h, C.byref(params), aw.AW_ORIENT_MASH, arr, len(pairs), 0, cb, None), "aw_align_pairs")
        if os.environ.get("AW_BENCH_TRACE"):
            print(f"[e2e] load {1e3 * (t_b - t_a):.1f} ms, align_pairs {1e3 * (time.perf_counter() - t_b):.1f} ms", file=sys.stderr)

    e2e_step()  # warm-up (allocations)
    acc.update(paf_bytes=0, n=0)
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        e2e_step()
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    paf_bytes_step = acc["paf_bytes"] // max(1, e2e_steps)
    h2d = sum(len(s) for s in seqs) + 8 * len(pairs) + sum(len(i) for i in ids)
    d2h = paf_bytes_step + 112 * len(pairs)
    ctx2.close()

    # ---- reduce over ranks: max time, summed work ----
    t = torch.tensor([ms, e2e_s * 1e3], dtype=torch.float64, device="cuda")
    w = torch.tensor([len(pairs), st["sum_block_len"], st["cells"]], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(w, op=dist.ReduceOp.SUM)
    ms_max, e2e_ms_max = float(t[0]), float(t[1])
    pairs_total, block_total, cells_total = float(w[0]), float(w[1]), float(w[2])

    if rank == 0:
        value = pairs_total * args.steps / (ms_max / 1e3)
        e2e_value = pairs_total * e2e_steps / (e2e_ms_max / 1e3)
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        peak = float(peaks.get("hbm_gbs", 6650.0))
        k_ms = sum(kernel_ms) / len(kernel_ms)
        algo = algorithmic_bytes(seqs, pairs, st["paf_bytes"])
        achieved = algo / (k_ms / 1e3) / 1e9
        traffic = None
        try:
            traffic = json.load(open(os.path.join(ROOT, "profiles", "ncu_summary.json"))).get("align_kernel", {}).get("dram_bytes_per_launch")
        except Exception:
            pass
        line = {
            "metric": "aligned pairs/s", "value": value, "unit": "pairs/s", "n_gpus": n_gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_max / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int16", "data": "synthetic",
            "config": {"workload": WORKLOAD, "pairs_per_gpu_per_step": B, "nseq": args.nseq, "partition": f"host LPT over {world} GPUs, no collective",
                       "l2": "per-launch working set (wavefront rings + history, >10 GB) exceeds the 126 MB L2; no explicit flush"},
            "gbp_per_s": block_total * args.steps / (ms_max / 1e3) / 1e9,
            "cells_per_s": cells_total * args.steps / (ms_max / 1e3),
            "clocks": {"sm_mhz": clocks["sm_mhz"], "sm_max_mhz": clocks["sm_max_mhz"], "reasons": clocks["reasons"]},
            "e2e": {"value": e2e_value, "unit": "pairs/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "steps": e2e_steps},
            "gpu_launches": launches_per_step * args.steps,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                         "kernel": "aw_align_kernel", "kernel_ms": k_ms, "algorithmic_bytes_per_launch": algo,
                         "peak_source": "measured (MEASURED_PEAKS.json)" if peaks else "fallback",
                         "note": "compulsory HBM traffic of this path is tiny (SURVEY 8d): the kernel is bound by integer issue + memory latency, see profiles/"},
        }
        if not args.no_cpu_baseline:
            import oracle_lib as O

            cores = os.cpu_count() or 1
            sample = pairs[: max(2, 8 * cores)]
            r = O.run_pairs(ids, seqs, sample, O.params(*SCORES), use_mash=True, threads=cores)
            line["cpu_baseline"] = {"value": len(sample) / r["seconds"], "unit": "pairs/s", "cores": cores, "kind": "port",
                                    "sample": f"first {len(sample)} pairs of rank 0's shard, {cores} threads, {r['seconds']:.1f} s"}
            # parity of the same pairs, GPU vs CPU restatement, outside every timed region
            gres = ctx.align_pairs(params, sample, orientation=aw.AW_ORIENT_MASH)
            same = sum(1 for g, c in zip(gres, r["paf"]) if g["paf"] == c)
            line["parity"] = {"pairs_checked": len(sample), "paf_identical": same, "scores_identical": sum(1 for g, s in zip(gres, r["scores"]) if g["score"] == s)}
            if same != len(sample):
                line["parity"]["FAILED"] = True
        print(json.dumps(line), flush=True)
    batch.close()
    ctx.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
