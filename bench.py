#!/usr/bin/env python
"""bench.py — clonotype-join throughput on B200 (BASELINE.json metric: exact-subclonotype pairs scored/sec).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

One "step" = one pass of the hot path (bucketer -> pair kernels -> forest -> union-find) over the whole
synthetic repertoire.  N=1 workload: BASELINE.json configs[1] (1M-cell heavy-tailed human BCR repertoire).
N>1 (weak scaling): the same clone-size law at N x 1M cells (N=8 ~ configs[4]); rank 0 generates the
repertoire, sub-buckets are assigned to ranks by a cost model (LPT), the heaviest cut into row ranges, and each
rank receives, uploads and joins only its share — no collective on the data path; the per-rank edge lists are
gathered GPU-to-GPU to rank 0 over NCCL and merged there (forest of the partial forests of split sub-buckets).

  value   pairs/s with inputs resident in HBM (ejb_run), device time from CUDA events on the launching stream
  e2e     pairs/s through ejb_join with HOST (pinned) buffers: H2D + run + D2H + EquivRel replay in the timed region
  roofline  dominant pair kernel vs the popc-pipe peak measured in the same job (SURVEY.md §8 d)
  cpu_baseline  the CPU oracle (port of the reference join) on a bounded bucket sample of the same workload

--impl reference times the reference's own CPU algorithm (the oracle port; the Rust reference cannot be
built in this image) on a bounded sample per step.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "exact-subclonotype pairs scored/sec"
CONFIG_INDEX = 1
WORKLOAD = "BASELINE configs[1]: synthetic 1M-cell human BCR repertoire, heavy-tailed clone sizes (Zipf 1.4, <=50k cells/clone)"


def bucket_table(inp):
    """Per-row bucket id (runs of equal lens, join.rs:68-88) and sub-bucket statistics computed with numpy."""
    a = inp.a
    n = inp.n_rows
    lens = a["row_len"].reshape(-1, 2)
    nch = a["row_nchains"]
    head = np.ones(n, dtype=bool)
    head[1:] = (lens[1:] != lens[:-1]).any(axis=1) | (nch[1:] != nch[:-1])
    bucket = np.cumsum(head) - 1
    return bucket, lens, nch


def algorithmic_popc(inp):
    """P1 = ceil(c_h/32)+ceil(c_l/32) popc per pair (stage 1);
    P2 = 3*(ceil(L_h/32)+ceil(L_l/32)) + 6 per stage-2 candidate.  Pair-weighted means over sub-buckets."""
    bucket, lens, nch = bucket_table(inp)
    c3 = inp.a["row_cdr3_len"].reshape(-1, 2)
    act = nch == 2
    key = np.stack([bucket[act], c3[act, 0], c3[act, 1], lens[act, 0], lens[act, 1]], axis=1)
    uniq, cnt = np.unique(key, axis=0, return_counts=True)
    pairs = cnt.astype(np.float64) * (cnt - 1) / 2
    p1 = np.ceil(uniq[:, 1] / 32) + np.ceil(uniq[:, 2] / 32)
    p2 = 3 * (np.ceil(uniq[:, 3] / 32) + np.ceil(uniq[:, 4] / 32)) + 6
    tot = max(pairs.sum(), 1.0)
    row_bytes = ((np.ceil(uniq[:, 1] / 4) + np.ceil(uniq[:, 2] / 4)) + 7 * 16 * (np.ceil(uniq[:, 3] / 128) + np.ceil(uniq[:, 4] / 128))) * cnt
    return float((p1 * pairs).sum() / tot), float((p2 * pairs).sum() / tot), float(pairs.sum()), float(row_bytes.sum())


def cpu_sample(inp, target_pairs=2.5e8, seed=0):
    """A bounded sample of the same workload for the CPU baseline: whole lens-buckets taken in a seeded
    shuffle until their pair count reaches target_pairs."""
    bucket, lens, nch = bucket_table(inp)
    nb = int(bucket[-1]) + 1
    sizes = np.bincount(bucket, minlength=nb).astype(np.int64)
    two = np.zeros(nb, dtype=bool)
    two[bucket[nch == 2]] = True
    order = np.random.default_rng(seed).permutation(nb)
    chosen, acc = [], 0.0
    for b in order:
        if not two[b] or sizes[b] < 2:
            continue
        p = sizes[b] * (sizes[b] - 1) / 2
        if p > target_pairs * 0.6:
            continue
        chosen.append(b)
        acc += p
        if acc >= target_pairs:
            break
    chosen = np.sort(np.array(chosen))
    rows = np.nonzero(np.isin(bucket, chosen))[0]
    desc = f"{len(chosen)} of {nb} lens-buckets of the workload (seeded shuffle, largest {int(sizes[chosen].max()) if len(chosen) else 0} rows), {len(rows)} rows"
    return inp.take_rows(rows), desc


class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index=0):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100", "-i", str(gpu_index)],
                                      stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.p is None:
            return out
        time.sleep(0.15)
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        self.f.seek(0)
        sm, mx, reasons = [], [], set()
        for line in self.f.read().splitlines():
            f = [x.strip() for x in line.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        try:
            os.unlink(self.f.name)
        except OSError:
            pass
        if sm:
            out.update(sm_mhz=statistics.median(sm), sm_max_mhz=max(mx), reasons=sorted(reasons), samples=len(sm))
        return out


def run_reference_arm(args, rank):
    """The reference's CPU join (oracle port), all host threads, bounded sample per step."""
    if rank != 0:
        return
    import enclone_ranger_b200 as E
    from oracle import pyoracle
    rep = E.generate(CONFIG_INDEX)
    inp = rep.to_join_input()
    rep.close()
    sample, desc = cpu_sample(inp, target_pairs=8e7)
    p = E.default_params(True)
    pyoracle.run(p, E.JoinInput().c_ptr(), threads=0)  # builds the Stirling table outside the timed region (start.rs:344 does too)
    times, pairs, threads = [], 0, 0
    for i in range(args.warmup + args.steps):
        r = pyoracle.run(p, sample.c_ptr(), threads=0)
        if i >= args.warmup:
            times.append(r["seconds_join"])
        pairs, threads = r["pairs_scored"], r["threads"]
    sec = sum(times) / max(len(times), 1)
    val = pairs / sec
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": "pairs/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u32 popc + f64 score",
        "data": "synthetic", "config": {"workload": WORKLOAD, "timed": "CPU restatement of join_exacts (oracle port of the Rust reference; rustc/cargo absent)",
                                          "sample": desc},
        "cpu_baseline": {"value": val, "unit": "pairs/s", "cores": threads, "kind": "port", "sample": desc},
        "e2e": {"value": val, "unit": "pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--cells", type=int, default=None, help="override the repertoire size (debug only; the number is then not the BASELINE config)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--verify", action="store_true", help="N>1 debug: rank 0 also joins the whole repertoire on its own GPU and compares the merged edge list and EquivRel")
    ap.add_argument("--split-above", type=float, default=0.5, help="N>1: split sub-buckets whose cost exceeds this fraction of a rank's share")
    ap.add_argument("--ascii", action="store_true", help="hand every contig over as ASCII (info[k].tigs) instead of 2-bit packed (tigsp)")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference_arm(args, rank)
        return
    args.warmup = max(args.warmup, 3)

    import torch
    import torch.distributed as dist
    import enclone_ranger_b200 as E
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the join has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    from enclone_ranger_b200 import multi
    base_cells = args.cells if args.cells is not None else 1000000
    t_gen = time.time()
    full_rows, row_exact_full = 0, None
    if world == 1:
        rep = E.generate(CONFIG_INDEX, n_cells=base_cells)
        inp = rep.to_join_input()      # numpy-owned copies (host buffers of the caller)
        rep.close()
        my_rows = None
        full_rows, row_exact_full = inp.n_rows, inp.a["row_exact"]
    else:
        parts = None
        if rank == 0:
            rep = E.generate(CONFIG_INDEX, n_cells=base_cells * world)
            full = rep.to_join_input()
            rep.close()
            full_rows, row_exact_full = full.n_rows, full.a["row_exact"].copy()
            parts = []
            # sample join on this rank's GPU: which large sub-buckets hold pairs that are scored in full and do not join
            probe_ctx = E.JoinContext(device=local_rank)
            live = multi.probe_live_fractions(full, probe_ctx, E.default_params(True))
            probe_ctx.close()
            wparts, split_rows = multi.partition_work(full, world, live_fraction=live, split_above=args.split_above)
            ncells_row = multi.ncells_per_row(full)
            for wp in wparts:
                sub, _ = full.subset_rows(wp.rows)
                parts.append((sub, wp.rows, wp.roles))
            n_split_rows = int(split_rows.sum())
            full_keep = full.to_packed() if args.verify else None
            del full
        inp, my_rows, my_roles = multi.scatter_parts(parts, dev)
        del parts
        fr = torch.tensor([full_rows], dtype=torch.int64, device=dev)
        dist.broadcast(fr, src=0)
        full_rows = int(fr.item())
    # host buffers as the Rust shim hands them over: contigs 2-bit packed (info[k].tigsp) where they are pure ACGT
    # without indel marks, ASCII otherwise
    raw_tig_bytes = int(inp.a["tig_bytes"].size)
    inp_ascii = inp
    if not args.ascii:
        inp = inp.to_packed()
    t_gen = time.time() - t_gen
    params = E.default_params(True)
    # the caller's host buffers live in page-locked memory (the e2e contract copies from pinned host memory): every array is
    # moved once, outside the timed regions, into ONE cudaHostAlloc'ed arena, the way the Rust shim flattens its records; the
    # library then uploads it with a single DMA (ejb_set_input_arena) instead of ~36 copies that each pay ~0.1 ms of setup
    views, total = {}, 0
    for name, arr in inp.a.items():
        total += (arr.nbytes + 255) & ~255
    arena = torch.empty(total + 256, dtype=torch.uint8).pin_memory()
    arena_np = arena.numpy()
    base_off = (-arena_np.ctypes.data) & 255           # 256-byte aligned start
    off = base_off
    for name, arr in inp.a.items():
        view = arena_np[off:off + arr.nbytes].view(arr.dtype).reshape(arr.shape)
        view[...] = arr
        views[name] = view
        off += (arr.nbytes + 255) & ~255
    pinned = [arena]
    arena_addr, arena_len = arena_np.ctypes.data + base_off, off - base_off
    inp = E.JoinInput(**views)     # same dtypes, contiguous: the struct points straight into the pinned buffers
    assert all(inp.a[n].ctypes.data == views[n].ctypes.data for n in views if views[n].size)
    ctx = E.JoinContext(device=local_rank)
    ctx.set_input_arena(arena_addr, arena_len)
    peaks = ctx.measure_peaks() if rank == 0 else None
    rows_dev = torch.from_numpy(my_rows).to(dev) if my_rows is not None else None
    if world > 1:
        ctx.set_row_roles(my_roles)   # row ranges of split sub-buckets (None: this rank holds whole sub-buckets only)
        split_dev = torch.from_numpy(split_rows).to(dev) if rank == 0 else None

    def gather_keys(key, aux):
        """per-rank 64-bit (k1<<32|k2) keys (global row ids) + (cd<<32|min_shares) -> rank 0 over NCCL; there: one device sort,
        then the edges of split sub-buckets (partial forests) are reduced to their forest and two-cell filtered"""
        if world == 1:
            return key
        n = torch.tensor([key.numel()], dtype=torch.int64, device=dev)
        sizes = [torch.zeros(1, dtype=torch.int64, device=dev) for _ in range(world)]
        dist.all_gather(sizes, n)
        sizes = [int(x.item()) for x in sizes]
        mx = max(max(sizes), 1)
        pad = torch.full((2, mx), -1, dtype=torch.int64, device=dev)
        pad[0, : key.numel()] = key
        pad[1, : key.numel()] = aux
        bufs = [torch.empty((2, mx), dtype=torch.int64, device=dev) for _ in range(world)] if rank == 0 else None
        dist.gather(pad, bufs, dst=0)
        if rank != 0:
            return None
        k, o = torch.sort(torch.cat([bufs[r][0, : sizes[r]] for r in range(world)]))
        if n_split_rows:
            a = torch.cat([bufs[r][1, : sizes[r]] for r in range(world)])[o]
            idx = torch.nonzero(split_dev[k >> 32]).squeeze(1)
            if idx.numel():
                ks, au = k[idx].cpu().numpy(), a[idx].cpu().numpy()
                keep = multi.merge_split_edges(full_rows, ks >> 32, ks & 0xffffffff, au >> 32, au & 0xffffffff, split_rows, ncells_row)
                mask = torch.ones(k.numel(), dtype=torch.bool, device=dev)
                mask[idx[torch.from_numpy(~keep).to(dev)]] = False
                k = k[mask]
        return k

    def edge_aux(cd, nrefs, shares):
        """(cd << 32) | min(shares): what the two-cell filter needs (join.rs:120-178)"""
        sh = shares.reshape(-1, 2).long()
        return (cd.long() << 32) | torch.where(nrefs == 2, torch.minimum(sh[:, 0], sh[:, 1]), sh[:, 0])

    # ---------------- device-resident arm: value ----------------
    ctx.upload(params, inp)

    gather_ms = []

    def step_resident():
        st = ctx.run()
        if world > 1:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            k1, k2 = ctx.device_edges()     # device pointers, no host round trip
            gather_keys((rows_dev[k1.long()] << 32) | rows_dev[k2.long()], edge_aux(*ctx.device_tuples()))
            e1.record()
            gather_ms.append((e0, e1))
        return st

    for _ in range(args.warmup):
        step_resident()
    barrier()
    gather_ms.clear()
    sampler = ClockSampler(local_rank) if rank == 0 else None
    dev_ms, wall0, launches, stats = [], time.perf_counter(), 0, None
    for _ in range(args.steps):
        stats = step_resident()
        dev_ms.append(stats["ms_device_total"])
        launches += stats["kernel_launches"]
    barrier()
    wall_ms = (time.perf_counter() - wall0) * 1e3 / args.steps
    clocks = sampler.stop() if sampler else None
    # per-step time on the device clock: the library's CUDA events around the join (its own stream) + at N>1 torch CUDA events
    # around the edge gather / merge (torch's stream; it starts after the join has completed); max over ranks below
    g_ms = [a.elapsed_time(b) for a, b in gather_ms]
    t_local = (sum(dev_ms) + sum(g_ms)) / len(dev_ms)
    pairs = stats["pairs_scored"]
    if world > 1:
        t = torch.tensor([t_local], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        t_step = float(t.item())
        ln = torch.tensor([launches, pairs], dtype=torch.int64, device=dev)
        dist.all_reduce(ln)
        launches, pairs = int(ln[0].item()), int(ln[1].item())
        keys = ("pairs_scored", "pairs_evaluated", "stage1_candidates", "stage1_ref_kills", "stage1_donor_kills", "stage2_survivors", "pass_edges", "rounds", "overflow_retries",
                "ms_pack", "ms_stage1", "ms_stage2", "ms_stage2a", "ms_forest", "ms_finalize")
        pr = [torch.zeros(2 + len(keys), dtype=torch.float64, device=dev) for _ in range(world)]
        dist.all_gather(pr, torch.tensor([wall_ms, sum(dev_ms) / len(dev_ms)] + [float(stats[k]) for k in keys], dtype=torch.float64, device=dev))
        per_rank = [dict({"wall_ms": round(float(x[0]), 2), "join_device_ms": round(float(x[1]), 2)}, **{k: round(float(v), 2) for k, v in zip(keys, x[2:].tolist())}) for x in pr]
    else:
        t_step = t_local
        per_rank = None

    # ---------------- end-to-end arm: host buffers through ejb_join ----------------
    last_merged = {}

    def step_e2e():
        res = ctx.join(params, inp, copy=False)   # zero-copy views into the library's pinned result arena
        if world > 1:
            key = torch.from_numpy(my_rows[res.edge_k1].astype(np.int64) << 32 | my_rows[res.edge_k2].astype(np.int64)).to(dev)
            aux = edge_aux(torch.from_numpy(np.ascontiguousarray(res.edge_cd)), torch.from_numpy(np.ascontiguousarray(res.edge_nrefs)),
                           torch.from_numpy(np.ascontiguousarray(res.edge_shares))).to(dev)
            merged = gather_keys(key, aux)
            if rank == 0:   # raw_joins of the whole repertoire + the EquivRel finish_join returns
                k = merged.cpu().numpy()
                last_merged["k"] = k
                last_merged["eq"] = E.equiv_replay(full_rows, (k >> 32).astype(np.int32), (k & 0xffffffff).astype(np.int32), row_exact_full)
        return res

    step_e2e()
    barrier()
    e2e_wall0 = time.perf_counter()
    res = None
    for _ in range(args.steps):
        res = step_e2e()
    barrier()
    e2e_ms = (time.perf_counter() - e2e_wall0) * 1e3 / args.steps
    if world > 1:
        t = torch.tensor([e2e_ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_ms = float(t.item())
    e2e_stats = res.stats
    if args.verify and world > 1 and rank == 0:
        vctx = E.JoinContext(device=local_rank)
        ref = vctx.join(params, full_keep)
        kref = ref.edge_k1.astype(np.int64) << 32 | ref.edge_k2.astype(np.int64)
        eq, lab = last_merged["eq"]
        ok = bool(np.array_equal(kref, last_merged["k"]) and np.array_equal(eq.x, ref.eq_x) and np.array_equal(eq.y, ref.eq_y) and np.array_equal(eq.z, ref.eq_z)
                  and np.array_equal(lab, ref.labels))
        print(f"[verify] {world}-rank merged result == single-GPU join of the whole repertoire: {ok} ({len(kref)} edges, {n_split_rows} rows in split sub-buckets)",
              file=sys.stderr, flush=True)
        vctx.close()
        if not ok:
            raise SystemExit("verification failed")

    if rank == 0:
        p1_avg, p2_avg, dom_pairs, packed_bytes = algorithmic_popc(inp_ascii)
        s1_ms, s2a_ms = stats["ms_stage1"], stats["ms_stage2a"]
        s1_ops = stats["pairs_evaluated"] * p1_avg
        s2_ops = stats["stage1_candidates"] * p2_avg
        kernels = {
            "k_stage1": {"ms_per_step": s1_ms, "launches_per_step": stats["stage1_launches"], "units": stats["pairs_evaluated"], "popc_per_unit": p1_avg,
                         "achieved_gpopc_s": s1_ops / max(s1_ms, 1e-9) / 1e6},
            "k_stage2a": {"ms_per_step": s2a_ms, "launches_per_step": stats["stage2a_launches"], "units": stats["stage1_candidates"], "popc_per_unit": p2_avg,
                          "achieved_gpopc_s": s2_ops / max(s2a_ms, 1e-9) / 1e6},
        }
        dom = "k_stage2a" if s2a_ms >= s1_ms else "k_stage1"
        peak_g = peaks["popc_per_s"] / 1e9
        hbm_peak = None
        try:
            hbm_peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
        except Exception:
            hbm_peak = 6650.0
        traffic = None   # DRAM bytes per launch of the dominant kernel, from the committed `ncu --set full` capture
        try:
            nt = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic_r1.json")))
            traffic = nt[dom]["dram_bytes_per_launch"]
        except Exception:
            pass
        alg_bytes = packed_bytes + 8.0 * stats["joins"] + 4.0 * inp.n_rows   # rank 0's share at N > 1
        line = {
            "metric": METRIC, "value": pairs / (t_step / 1e3), "unit": "pairs/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": t_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u32 popc + f64 score", "data": "synthetic",
            "config": {"workload": (WORKLOAD if world == 1 else f"configs[1] clone-size law at {world} x 1M cells, sub-buckets partitioned over {world} ranks (the heaviest cut into row ranges)") if args.cells is None else f"DEBUG {args.cells}-cell-per-GPU variant of configs[1]", "cells": int(base_cells * world), "rows": int(full_rows), "rows_rank0": int(inp.n_rows),
                       "pairs_scored": int(pairs), "pairs_lens_buckets": int(stats["pairs_lens"]), "sub_buckets": int(stats["n_subbuckets"]),
                       "parallelism": f"bucket-sharded x{world}, weak scaling (1M cells per GPU), {n_split_rows} rows in sub-buckets split by row ranges" if world > 1 else "single GPU",
                       "l2": f"inputs ({stats['h2d_bytes'] / 1e6:.0f} MB raw + packed row store) exceed the 126 MB L2; no flush needed",
                       "timing": "CUDA events on the library's launching stream (ejb_stats.ms_device_total)" if world == 1 else "device clock: CUDA events of the join on the library's stream + CUDA events of the NCCL edge gather and merge on torch's stream, max over ranks",
                       "wall_ms_per_step": wall_ms, "generate_s": t_gen, "per_rank": per_rank,
                       "contig_encoding": "ASCII" if args.ascii else f"2-bit packed where eligible ({raw_tig_bytes / 1e6:.0f} MB ASCII -> {inp.a['tig2_words'].nbytes / 1e6:.0f} MB packed + {inp.a['tig_bytes'].nbytes / 1e6:.0f} MB ASCII)"},
            "clocks": clocks,
            "e2e": {"value": pairs / (e2e_ms / 1e3), "unit": "pairs/s", "ms_per_step": e2e_ms, "h2d_bytes_per_step": int(e2e_stats["h2d_bytes"]),
                    "d2h_bytes_per_step": int(e2e_stats["d2h_bytes"]), "pinned_host_buffers": "one arena (ejb_set_input_arena)",
                    "breakdown_ms": {k: e2e_stats[k] for k in ("ms_h2d", "ms_device_total", "ms_d2h", "ms_replay")}},
            "gpu_launches": int(launches),
            "roofline": {"bound": "popc", "kernel": dom, "achieved": kernels[dom]["achieved_gpopc_s"], "peak": peak_g, "unit": "Gpopc/s",
                         "frac": kernels[dom]["achieved_gpopc_s"] / peak_g, "traffic": traffic,
                         "traffic_note": "dram__bytes_read.sum + dram__bytes_write.sum per launch, profiles/ncu_traffic_r1.json (ncu --set full, 1M-cell config[1])", "peak_source": "measured in this job (ejb_measure_peaks: dependency-free __popc loop)",
                         "lop3_peak_gops": peaks["lop3_per_s"] / 1e9, "kernels": kernels,
                         "whole_join": {"popc_per_pair": p1_avg + p2_avg * stats["stage1_candidates"] / max(pairs, 1),
                                        "roofline_pairs_per_s": peaks["popc_per_s"] / (p1_avg + p2_avg * stats["stage1_candidates"] / max(pairs, 1)),
                                        "frac": (pairs / (t_step / 1e3)) / (peaks["popc_per_s"] / (p1_avg + p2_avg * stats["stage1_candidates"] / max(pairs, 1)))},
                         "hbm": {"algorithmic_bytes": alg_bytes, "achieved_gbs": alg_bytes / (t_step / 1e3) / 1e9, "peak_gbs": hbm_peak}},
            "stage_ms": {k: stats[k] for k in ("ms_pack", "ms_stage1", "ms_stage2", "ms_stage2a", "ms_forest", "ms_finalize")},
            "counts": {k: int(stats[k]) for k in ("pairs_evaluated", "stage1_candidates", "stage1_ref_kills", "stage1_donor_kills", "stage2_two_ref", "stage2_memo_dead", "stage2_degr_dead", "stage2_slow", "stage2_survivors", "pass_edges", "forest_edges", "joins", "rounds")},
        }
        if world == 1 and not args.no_cpu_baseline:
            from oracle import pyoracle   # CPU baseline leg only: the checker timed as a baseline, never on the product path
            sample, desc = cpu_sample(inp_ascii)
            pyoracle.run(params, E.JoinInput().c_ptr(), threads=0)
            r = pyoracle.run(params, sample.c_ptr(), threads=0)
            line["cpu_baseline"] = {"value": r["pairs_scored"] / r["seconds_join"], "unit": "pairs/s", "cores": int(r["threads"]), "kind": "port",
                                    "sample": desc, "seconds": r["seconds_join"], "join_one_calls": int(r["join_one_calls"]), "pairs": int(r["pairs_scored"])}
        print(json.dumps(line))
    ctx.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
